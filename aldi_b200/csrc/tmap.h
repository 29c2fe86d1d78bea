// Host-side TMA tensor-map construction (driver API resolved at run time; no link-time libcuda dependency).
#pragma once
#include <cuda.h>
#include <stdint.h>

// bf16 tensor map with 128-byte swizzle and zero OOB fill. dims/box are innermost-first; `strides_bytes`
// has rank-1 entries (dims 1..rank-1).  Returns 0 or a negative ALDI error code.
int aldi_make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                        const uint64_t* strides_bytes, const uint32_t* box);

// same, with the shared-memory swizzle selectable: swizzle128 = 0 lays the box out linearly (rows of box[0] elements)
int aldi_make_tmap_bf16_sw(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                           const uint64_t* strides_bytes, const uint32_t* box, int swizzle128);
