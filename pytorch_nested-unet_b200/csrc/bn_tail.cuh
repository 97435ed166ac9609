// BatchNorm finalisation as the TAIL of the kernel that produced the per-CTA partial sums (conv epilogue statistics in forward,
// gradient-router sums in backward): the last CTA to finish (device-wide atomic counter, self-resetting) reduces the partial rows
// in a fixed order in fp64 — bitwise deterministic whichever CTA happens to be last — and emits the per-channel coefficients the
// next kernel consumes.  No separate finalize launch sits between a conv and its consumer any more (60 launches per step in round 1).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nunet {

struct BnFinalizeArgs {
    int mode;  // 0: forward statistics -> apply coefficients ; 1: backward sums -> dgamma/dbeta + g_y coefficients
    double count;
    const float* gamma;
    const float* beta;
    const float* conv_bias;
    float* running_mean;
    float* running_var;
    float momentum, eps;
    float* scale;
    float* shift;
    float* save_mean;
    float* save_invstd;
    // mode 1 (inputs: gamma, save_mean, save_invstd ; sums S1 = sum g_z, S2 = sum g_z*yhat)
    float* dgamma;
    float* dbeta;
    float* k1;
    float* k2;
    float* k3;
    float grad_scale;  // mode 1: multiplies dgamma / dbeta (1/world_size under SyncBN, see dist.py)
};

// SyncBN across the ranks of one node, INSIDE the tail (no collective launch): every rank's last CTA publishes its [2][C] fp64
// sums in its own peer-mapped slot, raises a flag in every peer's flag row (release, system scope), waits for the peers' flags
// (acquire) and sums all ranks' slots in rank order over NVLink peer loads — identical bits on every rank.  `seq` is a device
// counter of this slot, so replayed CUDA graphs keep exchanging fresh epochs; slots are double-buffered by epoch parity (a peer can
// only be one epoch ahead: it needs our flag of epoch e+1 before it may overwrite what we read at epoch e).
struct BnExchange {
    int world, rank;      // world <= 1: no exchange
    double* data[8];      // per rank: this slot's [2 parities][2][C] fp64 (peer-mapped)
    unsigned* flags[8];   // per rank: this slot's [8] flag words, word r written by rank r
    unsigned* seq;        // local epoch counter of this slot
};

struct BnTail {
    unsigned* counter;   // nullptr: no tail (the caller launches nunet_bn_finalize_train / nunet_bn_backward_reduce itself)
    double* sums_out;    // optional [2][C] fp64 (SyncBN exchange buffer)
    BnFinalizeArgs fin;  // fin.gamma == nullptr: only sums_out is produced
    BnExchange xch;      // in-kernel cross-rank reduction of the sums (fin.count is then the GLOBAL count)
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double* p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

// BatchNorm backward (nn.BatchNorm2d, train): dbeta = S1, dgamma = S2,
//   g_y = gamma*r*(g_z - S1/M - yhat*S2/M),  yhat = (y-mean)*r   =>   g_y = k1*g_z + k2*y + k3.
// The per-channel parameters a finalisation reads.  Loaded up front (bn_prefetch), before any of the result stores: the outputs are
// plain float* that may alias the inputs as far as the compiler knows, so loads placed after the first store were issued one by
// one — five dependent global-memory latencies (~9 k cycles of a forward tail, tools/prof_waits.py).
struct BnPre {
    float gamma, a, b, c, d;   // forward: a = beta, b = conv bias, c = running mean, d = running var ; backward: a = invstd, b = mean
};
__device__ __forceinline__ BnPre bn_prefetch(const BnFinalizeArgs& f, int ch) {
    BnPre q;
    q.gamma = f.gamma[ch];
    if (f.mode == 1) {
        q.a = f.save_invstd[ch];
        q.b = f.save_mean[ch];
        q.c = 0.f;
        q.d = 0.f;
    } else {
        q.a = f.beta[ch];
        q.b = f.conv_bias ? f.conv_bias[ch] : 0.f;
        q.c = f.running_mean ? f.running_mean[ch] : 0.f;
        q.d = f.running_mean ? f.running_var[ch] : 0.f;
    }
    return q;
}

__device__ __forceinline__ void bn_finalize_channel(const BnFinalizeArgs& a, int c, double S, double SS, const BnPre& q) {
    if (a.mode == 1) {
        // S = S1 = sum g_z, SS = S2 = sum g_z*yhat (header comment above)
        const double g = q.gamma, r = q.a, mu = q.b, M = a.count;
        a.dbeta[c] = (float)(S * a.grad_scale);
        a.dgamma[c] = (float)(SS * a.grad_scale);
        a.k1[c] = (float)(g * r);
        a.k2[c] = (float)(-g * r * r * SS / M);
        a.k3[c] = (float)(-g * r * S / M + g * r * r * mu * SS / M);
        return;
    }
    const double mean = S / a.count;
    double var = SS / a.count - mean * mean;
    if (var < 0.0) var = 0.0;
    const float invstd = (float)(1.0 / sqrt(var + (double)a.eps));
    const float sc = q.gamma * invstd;
    a.scale[c] = sc;
    a.shift[c] = q.a - (float)mean * sc;
    if (a.save_mean) a.save_mean[c] = (float)mean;
    if (a.save_invstd) a.save_invstd[c] = invstd;
    if (a.running_mean) {
        const double b = (double)q.b;  // the dropped conv bias only shifts the mean
        a.running_mean[c] = (float)((1.0 - a.momentum) * q.c + a.momentum * (mean + b));
        const double unbiased = a.count > 1.0 ? var * a.count / (a.count - 1.0) : var;
        a.running_var[c] = (float)((1.0 - a.momentum) * q.d + a.momentum * unbiased);
    }
}

__device__ __forceinline__ void bn_finalize_channel(const BnFinalizeArgs& a, int c, double S, double SS) {
    bn_finalize_channel(a, c, S, SS, bn_prefetch(a, c));
}

// Executed by the 256 threads `tid` = 0..255 of a CTA that share named barrier `bar_id`, AFTER they have written the CTA's own
// row(s) of part[rows][2][C] (C a multiple of 32, <= 512).  `scratch`: >= 8 KB + 16 B of shared memory that is free at this point.
// total_ctas = CTAs of the grid (every one of them must come through here exactly once).
// The reduction is latency-bound (one CTA, L2 loads): a thread owns four adjacent columns of the 2C-wide rows (one 16-byte load per
// row) and every L-th row, and keeps eight loads in flight; the L row lanes are then combined in a fixed order.
// Returns true in the (single) CTA that ran the finalisation, so that a caller can append further last-CTA work.
__device__ __forceinline__ bool bn_tail_run(const BnTail& t, const float* part, int rows, int C, unsigned total_ctas, int tid, int bar_id,
                                            void* scratch) {
    double* sh = reinterpret_cast<double*>(scratch);                      // [L][2C] partial sums (<= 1024 doubles)
    volatile int* flag = reinterpret_cast<volatile int*>(sh + 1024);
    __threadfence();                                                      // this CTA's partial row is visible device-wide ...
    asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");
    if (tid == 0) *flag = (atomicAdd(t.counter, 1u) == total_ctas - 1) ? 1 : 0;   // ... before it is counted
    asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");
    if (!*flag) return false;
    __threadfence();
    BnPre pre[2];                           // channels tid and tid + 256 (C <= 512): in flight under the partial-row loads below
#pragma unroll
    for (int k = 0; k < 2; ++k)
        if (t.fin.gamma && tid + k * 256 < C) pre[k] = bn_prefetch(t.fin, tid + k * 256);
    const int groups = C >> 1;              // 16-byte column groups per row (2C floats)
    const int L = 256 / groups;             // row lanes: 16 (C = 32) ... 1 (C = 512)
    const int g = tid % groups, lane = tid / groups;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    if (lane < L) {
        const float4* base = reinterpret_cast<const float4*>(part) + g;
        constexpr int U = 8;
        for (int r0 = lane; r0 < rows; r0 += L * U) {
            float4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int r = r0 + u * L;
                v[u] = (r < rows) ? __ldcg(base + (size_t)r * groups) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) { acc[0] += (double)v[u].x; acc[1] += (double)v[u].y; acc[2] += (double)v[u].z; acc[3] += (double)v[u].w; }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) sh[(size_t)lane * 2 * C + g * 4 + e] = acc[e];
    }
    asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");
    const bool exchange = t.xch.world > 1;
    unsigned epoch = 0;
    if (exchange) {
        epoch = *t.xch.seq + 1;                              // same value on every rank: all ranks run the same launches
        double* mine = t.xch.data[t.xch.rank] + (size_t)(epoch & 1u) * 2 * C;
        for (int c = tid; c < C; c += 256) {
            double S = 0.0, SS = 0.0;
            for (int l = 0; l < L; ++l) { S += sh[(size_t)l * 2 * C + c]; SS += sh[(size_t)l * 2 * C + C + c]; }
            mine[c] = S;
            mine[C + c] = SS;
        }
        __threadfence_system();
        asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");
        if (tid < t.xch.world && tid != t.xch.rank) {
            st_release_sys(t.xch.flags[tid] + t.xch.rank, epoch);           // "rank `rank` has published epoch `epoch`"
            unsigned spins = 0;
            while ((int)(ld_acquire_sys(t.xch.flags[t.xch.rank] + tid) - epoch) < 0) {
                if (++spins > (1u << 28)) { __trap(); }                      // a missing peer must not hang the GPU
            }
        }
        asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int c = tid + k * 256;
        if (c >= C) break;
        double S = 0.0, SS = 0.0;
        if (exchange) {
            for (int q = 0; q < t.xch.world; ++q) {          // rank order: bitwise identical totals on every rank
                const double* src = t.xch.data[q] + (size_t)(epoch & 1u) * 2 * C;
                S += ld_relaxed_sys_f64(src + c);
                SS += ld_relaxed_sys_f64(src + C + c);
            }
        } else {
            for (int l = 0; l < L; ++l) { S += sh[(size_t)l * 2 * C + c]; SS += sh[(size_t)l * 2 * C + C + c]; }
        }
        if (t.sums_out) { t.sums_out[c] = S; t.sums_out[C + c] = SS; }
        if (t.fin.gamma) bn_finalize_channel(t.fin, c, S, SS, pre[k]);
    }
    if (tid == 0) {
        *t.counter = 0;   // self-resetting for the next launch
        if (exchange) *t.xch.seq = epoch;
    }
    return true;
}

}  // namespace nunet
