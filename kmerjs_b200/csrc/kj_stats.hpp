// kj_stats.hpp -- exact-decimal statistics (kj_stats.cpp), used by kj_score.cu to finish rows
#pragma once
#include <stdint.h>
#include <string>
#include "../../include/kmerjs_b200.h"

bool kj_exact_zscore(int rm, uint64_t r1, uint64_t n1, uint64_t r2, uint64_t n2, double *z,
                     std::string *z_text);
bool kj_exact_fastp_text(const char *z_text, double *p);
bool kj_exact_row(int rm, uint64_t uscore, uint64_t tscore, uint64_t uscore0, uint64_t tscore0,
                  uint64_t lengths, uint64_t ulength, uint64_t hits, uint64_t kmer_map_size,
                  uint64_t summary_templates, uint64_t summary_unique_lens, kj_row *out,
                  int *accepted);
