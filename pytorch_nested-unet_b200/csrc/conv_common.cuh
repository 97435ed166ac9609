// Shared declarations of the conv kernels (kernel parameter blocks mirror include/nunet_b200.h descriptors).
#pragma once
#include <cuda.h>
#include "ptx.cuh"
#include "../../include/nunet_b200.h"

namespace nunet {

struct ConvSrc {
    const void* data;
    const void* data2;
    const float* scale;
    const float* shift;
    const float* k2;
    int C;
    int kind;
};

struct ConvParams {
    CUtensorMap tmap[NUNET_MAX_SRC];  // TMA mode only: one tiled map per pre-activated source (box = 32 ch x (W+2) px, swizzle 64B)
    int N, H, W;
    int Wp, Hp1;  // W+2, H+1
    int Q;        // padded positions: (N*(H+1)+1)*(W+2)
    int T, NT, CK, S, P, nchunks, cin_pad, Cout, num_src, tmem_cols, ntg;
    int res_b;          // 1: single K chunk, the weight block is fetched once per CTA and stays resident
    int a_stage_bytes;  // TMA mode: bytes of one A stage ((P + Wp) rows of 64 B, rounded to 1 KB)
    ConvSrc src[NUNET_MAX_SRC];
    const void* wpack;
    void* out;
    float* stat_part;
};

}  // namespace nunet
