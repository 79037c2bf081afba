// node/index.js -- the reference's module surface (lib/kmers.js, lib/kmerFinderClient.js, README.md)
// re-exposed over the N-API addon.  Same names, argument order and error texts; see INTEGRATION.md.
// Not runnable in this repository's images (no Node.js); kmerjs_b200/*.py is the tested twin.
'use strict';
const native = require('./build/Release/kmerjs_b200.node');
const fs = require('fs');
const EventEmitter = require('events');

let ctx = null;
function context() { if (!ctx) ctx = native.init(Number(process.env.LOCAL_RANK || 0)); return ctx; }

const complementMap = new Map([['A', 'T'], ['T', 'A'], ['G', 'C'], ['C', 'G']]);   // lib/kmers.js:12-17
function complement(s) {                                                            // lib/kmers.js:31-38
  return s.replace(/[ATGC]/g, (m) => complementMap.get(m)).split('').reverse().join('');
}
function mapToJSON(m) { const o = Object.create(null); for (const [k, v] of m) o[k] = v; return o; }
function objectToMap(o) { return new Map(Object.entries(o)); }

class KmerJS {                                                                      // lib/kmers.js:56-186
  constructor(fastq = '', preffix = 'ATGAC', length = 16, step = 1, coverage = 1, progress = true, env = 'node') {
    Object.assign(this, { fastq, preffix, kmerLength: length, step, coverage, progress, env });
    this.kmerMap = new Map(); this.kmerMapSize = 0; this.lines = 0; this.bytesRead = 0;
  }
  readFile() {
    const event = new EventEmitter();
    const promise = native.countFile(context(), this.fastq, this.preffix, this.kmerLength, this.step)
      .then((counts) => {
        const e = native.countsExport(counts);
        const m = new Map();
        for (let i = 0; i < e.keys.length; i++) m.set(e.keys[i], e.counts[i]);   // first-insertion order
        m._counts = counts;                       // device table, consumed by findFirstMatch
        this.kmerMap = m; this.kmerMapSize = m.size; this.lines = e.lines; this.bytesRead = e.bytesRead;
        event.emit('progress', { percentage: 100, transferred: e.bytesRead });
        return m;
      });
    return { promise, event };
  }
}

function kmerjs(fastqPath, prefix = 'ATGAC', k = 16, step = 1, output) {           // README.md:12-16
  const job = new KmerJS(fastqPath, prefix, k, step, 1, false);
  return job.readFile().promise.then((m) => {
    if (output) {                                                                   // lib/index.js:381-388
      let s = '{\n'; for (const [key, v] of m) s += `${key}: ${v},`; fs.writeFileSync(output, s + '}\n');
    }
    return m;
  });
}

module.exports = { kmerjs, default: kmerjs, KmerJS, complement, complementMap, mapToJSON, objectToMap };
