"""oracle/synth_ref.py (numpy) against the device generator (kj_synth_kernel): byte for byte, so that the CPU legs of
bench.py see exactly the reads the GPU arm counts."""
import numpy as np
import pytest

import synth_ref

pytestmark = pytest.mark.gpu


def test_numpy_generator_equals_device_generator():
    from kmerjs_b200 import synth
    for first in (0, 12345, 3 * 10_000_000):
        w = synth.Workload(n_reads=3000, genome_len=200_000, seed=0x6B6D6572, first_read=first)
        g = synth_ref.genome(0x6B6D6572, 200_000)
        assert g.tobytes() == w.genome_host()
        assert synth_ref.fastq(0x6B6D6572, 3000, g, first_read=first).tobytes() == w.host_bytes()
    w = synth.Workload(n_reads=500, genome_len=50_000, seed=77, sub_rate=0.05, n_rate=0.01, lead_n_rate=0.3)
    g = synth_ref.genome(77, 50_000)
    assert synth_ref.fastq(77, 500, g, sub_rate=0.05, n_rate=0.01, lead_n_rate=0.3).tobytes() == w.host_bytes()
