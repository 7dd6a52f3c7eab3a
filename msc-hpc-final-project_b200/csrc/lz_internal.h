// lz_internal.h — declarations shared by the translation units of liblzb200.so (not part of the ABI).
#pragma once
#include <stdint.h>
#include <stddef.h>
#include "../../include/lz.h"

int lz_fail(int code, const char* fmt, ...);   // records the message for lz_last_error(), returns code

// Host CSR builder used by lz_graph_generate_host (lz_host.cc).
int lz_build_csr_from_keys(uint64_t n, uint64_t* keys, uint64_t nkeys, uint64_t* nnz_out,
                           uint32_t** row_offset_out, uint32_t** col_idx_out);
