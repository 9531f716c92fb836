// tc_gemm.cuh - bf16 x bf16 -> fp32 GEMM on tcgen05 tensor cores, operands fed by TMA.
//
//   D[z][m, n] = sum_k A[a_row_off(z) + m, k] * B[b_row_off(z) + n, k + b_k_off(z)]   (+ bias[n])
//
// Both operands are K-major bf16 matrices described by 2-D tensor maps (128B swizzle, box 64 x 128).
// One 128 x BN output tile per CTA (BN = 128), fp32 accumulator in TMEM, 3-stage smem ring:
//   warp 0   TMA producer (one elected lane)
//   warp 1   TMEM allocation + MMA issue (one elected lane, tcgen05.mma cta_group::1, M=128, N=BN, K=16)
//   warps 2-5 epilogue: tcgen05.ld 32 lanes x 32 columns -> bias -> fp32 / bf16 store or fp32 atomic add
// grid.z enumerates (batch z, split-K slice).  Out-of-range rows / K are zero-filled by TMA; the epilogue
// guards m < M, n < N.  A negative or past-the-end K coordinate (b_k_off) reads zeros - used to express
// the time-shifted h_{t-1} operand of dW_hh without materialising it.
#pragma once
#include "tc_common.cuh"
#include <cstring>

namespace tcg {

constexpr int BM = 128, BN = 128, BK = 64, MAX_STAGES = 4;   // 3 x 32 KB: two CTAs per SM, so one tile's
                                                              // epilogue overlaps the other's main loop
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;
static inline int smem_bytes_for(int stages, int stage_out = 65536) {
    const int ring = stages * (A_BYTES + B_BYTES);
    return (ring < stage_out ? stage_out : ring) + 1024 /*align slack*/ + 256 /*barriers*/;
}
constexpr int THREADS = 192;

enum { OUT_F32 = 0, OUT_BF16 = 1, OUT_ATOMIC_F32 = 2, OUT_SCAN_BF16 = 3, OUT_SCAN_F32 = 4 };

// "scan-private" blocked output (modes OUT_SCAN_*): rows m = (d, g, unit), columns n = (t, b) are scattered
// so that every thread of the recurrence kernel finds the values of one time step in one contiguous run:
//   elem(m, n) = (((((d*ntiles + tile)*T + t)*CS + c)*G + g)*256 + tid)*8 + i        (gate-major inside a block)
//   tile = b/16, c = unit/128, tid = ((unit%128)/32 + 4*((b%16)/8))*32 + unit%32, i = b%8
// so one (direction, tile, step, CTA) block is G*4 KB contiguous (one bulk copy for the scan kernel) and a warp of
// this epilogue (32 consecutive units, fixed b-run) writes 512 contiguous bytes.
struct ScanBlk { int T, B, H, G; int U, NB; };   // U units per CTA (128: bf16 scans, 64: x3 scans), NB batch rows per tile (16 / 32)
// general form of the mapping above: tile = b/NB, c = unit/U, tid = unit%U + U*((b%NB)/8), i = b%8

struct Params {
    int M, N, K;              // K = full reduction length (split across splitk slices)
    int batch, splitk;
    int mode;
    void* C; int64_t ldc;     // row stride in elements
    int64_t zC;               // element offset of batch z in C
    int a_row_off[4], b_row_off[4], b_k_off[4];
    int a_mn, b_mn;           // 1: operand stored [K rows][MN contiguous] (MN-major), tensor map box 64(MN) x 64(K)
    const float* bias;        // per output column n - or per row m when bias_per_row - (nullable), batch stride zBias
    int64_t zBias;
    int bias_per_row;
    ScanBlk blk;              // OUT_SCAN_* geometry
    int m_fast;               // rasterisation: 1 = consecutive CTAs walk m-tiles first (B tile shared through L2)
    int stages;               // smem ring depth (1..4), chosen per problem: shallow rings let 3-4 CTAs share an SM
    int tma_store;            // 1: epilogue stages the tile in smem and writes it with TMA (store / reduce-add)
    int b_win;                // > 0: the B operand is a set of stride-1 sliding windows over a chunk [rows][F] (the collated batch
                              //      x[b,t,:] = chunk[b + t,:] is never materialised): logical row r = t*b_win + b (K-major B: the
                              //      N index, MN-major B: the K index) lives at chunk row r / b_win + r % b_win.  A tile's 128 (64)
                              //      rows must share one t: b_win % 128 (64) == 0.
    int b_win_rows;           // host side: rows of that chunk (extent of the tensor map of a windowed MN-major B operand)
    int nsplit;               // 1: plain bf16 product; 3: fp32-class product of split operands, A_hi B_hi + A_hi B_lo + A_lo B_hi
                              //    (second pair of tensor maps = the low parts; three ring slots per k-block)
    unsigned int* dbg;        // watchdog record (nullable)
};

__global__ void __launch_bounds__(THREADS, 4)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmAlo, const __grid_constant__ CUtensorMap tmBlo,
            const __grid_constant__ CUtensorMap tmC0, const __grid_constant__ CUtensorMap tmC1, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int STAGES = p.stages;
    // the TMA-store epilogue needs up to 64 KB of staging (fp32 tile); with shallow rings it gets its own space
    const int ring_bytes = STAGES * (A_BYTES + B_BYTES);
    const int stage_out = p.tma_store ? ((p.mode == OUT_BF16 || p.mode == OUT_SCAN_BF16) ? 32768 : 65536) : 0;
    const int data_bytes = ring_bytes < stage_out ? stage_out : ring_bytes;
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + data_bytes);
    uint64_t* full = bars;                 // [STAGES] TMA -> MMA
    uint64_t* empty = bars + 4;            // [STAGES] MMA -> TMA
    uint64_t* accum = bars + 8;            // MMA -> epilogue
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int z = blockIdx.z / p.splitk, ks = blockIdx.z % p.splitk;
    const int m0 = (p.m_fast ? blockIdx.x : blockIdx.y) * BM, n0 = (p.m_fast ? blockIdx.y : blockIdx.x) * BN;
    const int kblocks_total = (p.K + BK - 1) / BK;
    const int kb_per = (kblocks_total + p.splitk - 1) / p.splitk;
    const int kb0 = ks * kb_per;
    const int nkb = max(0, min(kblocks_total, kb0 + kb_per) - kb0);

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
        tc::mbar_init(accum, 1);
        tc::fence_mbar_init();
    }
    if (warp == 1) tc::tmem_alloc(tmem_slot, BN);
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (tc::elect_one()) {
            tc::tma_prefetch_desc(&tmA);
            tc::tma_prefetch_desc(&tmB);
            if (p.nsplit > 1) { tc::tma_prefetch_desc(&tmAlo); tc::tma_prefetch_desc(&tmBlo); }
            const int nit = nkb * p.nsplit;
            int combo = 0, kbi = kb0, s = 0;                        // advanced incrementally: no divisions on the issue path
            uint32_t ph = 0;
            for (int i = 0; i < nit; ++i) {
                if (!tc::mbar_wait(&empty[s], ph ^ 1, p.dbg, 0x100 + s)) break;
                tc::mbar_arrive_expect_tx(&full[s], A_BYTES + B_BYTES);
                // combo 0: hi x hi, 1: hi x lo, 2: lo x hi
                const int k = kbi * BK;
                // (each call names its tensor map directly: a run-time selected pointer makes the compiler copy the 128-byte
                // descriptor to local memory)
                const bool a_lo = combo == 2, b_lo = combo == 1;
                uint8_t* da = sA + s * A_BYTES;
                uint8_t* db = sB + s * B_BYTES;
                const int ar = p.a_row_off[z] + m0;
                int br = p.b_row_off[z] + n0, kb = k + p.b_k_off[z];
                if (p.b_win) { if (p.b_mn) kb = kb / p.b_win + kb % p.b_win; else br = p.b_row_off[z] + n0 / p.b_win + n0 % p.b_win; }
                if (p.a_mn) {        // two boxes of 64 (MN) x 64 (K rows)
                    if (a_lo) { tc::tma_load_2d(da, &tmAlo, &full[s], ar, k); tc::tma_load_2d(da + A_BYTES / 2, &tmAlo, &full[s], ar + 64, k); }
                    else { tc::tma_load_2d(da, &tmA, &full[s], ar, k); tc::tma_load_2d(da + A_BYTES / 2, &tmA, &full[s], ar + 64, k); }
                } else {
                    if (a_lo) tc::tma_load_2d(da, &tmAlo, &full[s], k, ar);
                    else tc::tma_load_2d(da, &tmA, &full[s], k, ar);
                }
                if (p.b_mn) {
                    if (b_lo) { tc::tma_load_2d(db, &tmBlo, &full[s], br, kb); tc::tma_load_2d(db + B_BYTES / 2, &tmBlo, &full[s], br + 64, kb); }
                    else { tc::tma_load_2d(db, &tmB, &full[s], br, kb); tc::tma_load_2d(db + B_BYTES / 2, &tmB, &full[s], br + 64, kb); }
                } else {
                    if (b_lo) tc::tma_load_2d(db, &tmBlo, &full[s], kb, br);
                    else tc::tma_load_2d(db, &tmB, &full[s], kb, br);
                }
                if (++combo == p.nsplit) { combo = 0; ++kbi; }
                if (++s == STAGES) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (tc::elect_one()) {
            const uint32_t idesc = tc::umma_idesc_bf16(BM, BN, (uint32_t)p.a_mn, (uint32_t)p.b_mn);
            const int nit = nkb * p.nsplit;
            int s = 0;
            uint32_t ph = 0;
            for (int i = 0; i < nit; ++i, s = (s + 1 == STAGES ? 0 : s + 1), ph ^= (s == 0 ? 1u : 0u)) {
                if (!tc::mbar_wait(&full[s], ph, p.dbg, 0x200 + s)) break;
                tc::tcgen05_fence_after();
                const uint32_t aa = tc::smem_u32(sA + s * A_BYTES), ab = tc::smem_u32(sB + s * B_BYTES);
                const uint64_t da = p.a_mn ? tc::umma_desc_mn_sw128(aa, A_BYTES / 2) : tc::umma_desc_k_sw128(aa);
                const uint64_t db = p.b_mn ? tc::umma_desc_mn_sw128(ab, B_BYTES / 2) : tc::umma_desc_k_sw128(ab);
                // per K=16 step: K-major advances 32 B inside the swizzled row; MN-major advances 16 K-rows = 2048 B
                const uint64_t sa = p.a_mn ? 128 : 2, sb = p.b_mn ? 128 : 2;
#pragma unroll
                for (int kk = 0; kk < BK / 16; ++kk)
                    tc::umma_bf16(tmem, da + sa * kk, db + sb * kk, idesc, (i | kk) ? 1u : 0u);
                tc::umma_commit(&empty[s]);               // frees the smem slot when these MMAs retire
            }
            tc::umma_commit(accum);
        }
    } else {
        // epilogue warps 2..5 -> TMEM lane quarter = warp % 4
        const int q = warp & 3;
        const int m = m0 + q * 32 + lane;
        bool ok = tc::mbar_wait(accum, 0, p.dbg, 0x300);
        tc::tcgen05_fence_after();
        const float* bias = p.bias ? p.bias + z * p.zBias : nullptr;
        const float brow = (bias && p.bias_per_row && m < p.M) ? bias[m] : 0.f;
        if (ok && p.mode >= OUT_SCAN_BF16 && p.tma_store) {
            // ---- blocked ("scan-private") output through smem + 1-D bulk stores: for every 8-column run the 128 rows of
            // this tile form one contiguous block [128 units][8] in the destination (see ScanBlk), staged at the same
            // shape in smem (16-byte / 32-byte per thread: conflict-free) and written by cp.async.bulk.
            const int ml = q * 32 + lane;
            const bool bf = p.mode == OUT_SCAN_BF16;
            const int run_bytes = bf ? 2048 : 4096;
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                uint32_t v[32];
                if (nkb > 0) { tc::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + c * 32, v); tc::tmem_ld_wait(); }
                else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = 0u;
                }
                const int nb = n0 + c * 32;
#pragma unroll
                for (int i8 = 0; i8 < 32; i8 += 8) {
                    float f[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        f[j] = __uint_as_float(v[i8 + j]) + (bias ? (p.bias_per_row ? brow : ((nb + i8 + j < p.N) ? bias[nb + i8 + j] : 0.f)) : 0.f);
                    uint8_t* dst = smem + (size_t)(c * 4 + (i8 >> 3)) * run_bytes;
                    if (bf) {
                        uint32_t w[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
                            w[j] = *reinterpret_cast<uint32_t*>(&h2);
                        }
                        tc::sts_u4(tc::smem_u32(dst) + ml * 16, make_uint4(w[0], w[1], w[2], w[3]));      // explicit st.shared (not generic)
                    } else {
                        tc::sts_f4(tc::smem_u32(dst) + ml * 32, make_float4(f[0], f[1], f[2], f[3]));
                        tc::sts_f4(tc::smem_u32(dst) + ml * 32 + 16, make_float4(f[4], f[5], f[6], f[7]));
                    }
                }
            }
            tc::fence_proxy_async_smem();
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (warp == 2 && tc::elect_one()) {
                const int H = p.blk.H, G = p.blk.G, Bb = p.blk.B, Tt = p.blk.T, U = p.blk.U;
                const int sh = p.blk.NB == 32 ? 5 : 4;              // batch rows per tile: 16 or 32
                const int CSs = H / U, ntl = Bb >> sh, nh = BM / U;
                const size_t es = bf ? 2 : 4;
                // the tile's 128 rows = BM/U runs of U units of one (d, gate, cta): their block coordinates are fixed for the tile
                size_t row_part[2]; size_t gate_part[2];
                for (int hh = 0; hh < nh; ++hh) {
                    const int mr = m0 + hh * U;
                    row_part[hh] = (size_t)(mr / (G * H)) * ntl;
                    gate_part[hh] = (size_t)((mr % H) / U) * G + (size_t)((mr / H) % G);
                }
                int t_ = n0 / Bb, b = n0 % Bb;                      // advanced incrementally: 8 columns per run
                for (int r = 0; r < BN / 8; ++r) {
                    if (n0 + r * 8 >= p.N) break;
                    const int tile_ = b >> sh, cbg = (b & ((1 << sh) - 1)) >> 3;
                    for (int hh = 0; hh < nh; ++hh) {
                        const size_t e = ((((row_part[hh] + tile_) * Tt + t_) * CSs * G + gate_part[hh]) * 256 + (size_t)cbg * U) * 8;
                        tc::bulk_s2g(reinterpret_cast<uint8_t*>(p.C) + e * es, smem + (size_t)r * run_bytes + (size_t)hh * U * 8 * es,
                                     (uint32_t)(U * 8 * es));
                    }
                    b += 8;
                    if (b >= Bb) { b -= Bb; ++t_; }
                }
                tc::tma_store_commit();
                tc::tma_store_wait_read();     // smem may be released once it has been read; the writes complete on their own
            }
        } else if (ok && p.tma_store) {
            // ---- staged epilogue: TMEM -> registers -> 128B-swizzled smem boxes -> TMA store / reduce-add.
            // The pipeline stages are free (every MMA has retired), so they serve as the staging buffer:
            // bf16: 2 boxes of [128 rows x 64 cols], fp32: 4 boxes of [128 rows x 32 cols], 16 KB each.
            const int ml = q * 32 + lane;
            const uint32_t sw = (uint32_t)(ml & 7);
            const bool is_bf16 = p.mode == OUT_BF16;
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                uint32_t v[32];
                if (nkb > 0) {
                    tc::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + c * 32, v);
                    tc::tmem_ld_wait();
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = 0u;
                }
                const int nb = n0 + c * 32;
                if (bias && ks == 0) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float bb = p.bias_per_row ? brow : ((nb + i < p.N) ? bias[nb + i] : 0.f);
                        v[i] = __float_as_uint(__uint_as_float(v[i]) + bb);
                    }
                }
                if (is_bf16) {
                    uint8_t* box = smem + (size_t)(c >> 1) * 16384 + (size_t)ml * 128;
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) {
                        uint32_t w[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(v[k4 * 8 + 2 * j]), __uint_as_float(v[k4 * 8 + 2 * j + 1]));
                            w[j] = *reinterpret_cast<uint32_t*>(&h2);
                        }
                        const uint32_t chunk = ((uint32_t)((c & 1) * 4 + k4)) ^ sw;
                        tc::sts_u4(tc::smem_u32(box) + chunk * 16, make_uint4(w[0], w[1], w[2], w[3]));
                    }
                } else {
                    uint8_t* box = smem + (size_t)c * 16384 + (size_t)ml * 128;
#pragma unroll
                    for (int k4 = 0; k4 < 8; ++k4) {
                        const uint32_t chunk = ((uint32_t)k4) ^ sw;
                        tc::sts_u4(tc::smem_u32(box) + chunk * 16, make_uint4(v[k4 * 4], v[k4 * 4 + 1], v[k4 * 4 + 2], v[k4 * 4 + 3]));
                    }
                }
            }
            tc::fence_proxy_async_smem();
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (warp == 2 && tc::elect_one()) {
                const int nboxes = is_bf16 ? BN / 64 : BN / 32;
                const int bw = is_bf16 ? 64 : 32;
                for (int b = 0; b < nboxes; ++b) {
                    if (n0 + b * bw >= p.N) break;
                    // (the map is named directly in each call: a run-time selected pointer costs a 128-byte local copy)
                    if (p.mode == OUT_ATOMIC_F32) {
                        if (z == 0) tc::tma_reduce_add_2d(&tmC0, smem + (size_t)b * 16384, n0 + b * bw, m0);
                        else tc::tma_reduce_add_2d(&tmC1, smem + (size_t)b * 16384, n0 + b * bw, m0);
                    } else {
                        if (z == 0) tc::tma_store_2d(&tmC0, smem + (size_t)b * 16384, n0 + b * bw, m0);
                        else tc::tma_store_2d(&tmC1, smem + (size_t)b * 16384, n0 + b * bw, m0);
                    }
                }
                tc::tma_store_commit();
                tc::tma_store_wait_read();     // smem may be released once it has been read; the writes complete on their own
            }
        } else if (ok) {
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                uint32_t v[32];
                if (nkb > 0) {
                    tc::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + c * 32, v);
                    tc::tmem_ld_wait();
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = 0u;
                }
                const int nb = n0 + c * 32;
                if (m < p.M && nb < p.N && p.mode >= OUT_SCAN_BF16) {
                    // blocked store: 8 consecutive b (one thread-run of the scan kernel) = one 16/32-byte store
                    const int H = p.blk.H, G = p.blk.G, Bb = p.blk.B, Tt = p.blk.T, U = p.blk.U, NBt = p.blk.NB;
                    const int dd = m / (G * H), gg = (m / H) % G, unit = m % H;
                    const int CSs = H / U, cc = unit / U, ju = unit % U;
                    const int ntl = Bb / NBt;
#pragma unroll
                    for (int i8 = 0; i8 < 32; i8 += 8) {
                        const int n = nb + i8;
                        if (n >= p.N) break;
                        const int t = n / Bb, b = n % Bb;
                        const int tile = b / NBt;
                        const int tid = ju + U * ((b % NBt) >> 3);
                        const size_t e = ((((((size_t)dd * ntl + tile) * Tt + t) * CSs + cc) * G + gg) * 256 + tid) * 8;
                        float f[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[i8 + j]) + (bias ? (p.bias_per_row ? brow : bias[n + j]) : 0.f);
                        if (p.mode == OUT_SCAN_BF16) {
                            uint32_t w[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
                                w[j] = *reinterpret_cast<uint32_t*>(&h2);
                            }
                            *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.C) + e) = make_uint4(w[0], w[1], w[2], w[3]);
                        } else {
                            float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.C) + e);
                            o[0] = make_float4(f[0], f[1], f[2], f[3]);
                            o[1] = make_float4(f[4], f[5], f[6], f[7]);
                        }
                    }
                } else if (m < p.M && nb < p.N) {
                    if (p.mode == OUT_BF16) {
                        __nv_bfloat16* crow = reinterpret_cast<__nv_bfloat16*>(p.C) + z * p.zC + (int64_t)m * p.ldc + nb;
                        if (nb + 32 <= p.N && (p.ldc % 8 == 0)) {
#pragma unroll
                            for (int i = 0; i < 32; i += 8) {
                                uint32_t w[4];
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    float a = __uint_as_float(v[i + 2 * j]), b = __uint_as_float(v[i + 2 * j + 1]);
                                    if (bias) {
                                        if (p.bias_per_row) { a += brow; b += brow; }
                                        else { a += bias[nb + i + 2 * j]; b += bias[nb + i + 2 * j + 1]; }
                                    }
                                    __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
                                    w[j] = *reinterpret_cast<uint32_t*>(&h2);
                                }
                                *reinterpret_cast<uint4*>(crow + i) = make_uint4(w[0], w[1], w[2], w[3]);
                            }
                        } else {
                            for (int i = 0; i < 32 && nb + i < p.N; ++i) {
                                float a = __uint_as_float(v[i]);
                                if (bias) a += p.bias_per_row ? brow : bias[nb + i];
                                crow[i] = __float2bfloat16(a);
                            }
                        }
                    } else {
                        float* crow = reinterpret_cast<float*>(p.C) + z * p.zC + (int64_t)m * p.ldc + nb;
                        for (int i = 0; i < 32 && nb + i < p.N; ++i) {
                            float a = __uint_as_float(v[i]);
                            if (bias && ks == 0) a += p.bias_per_row ? brow : bias[nb + i];
                            if (p.mode == OUT_ATOMIC_F32) atomicAdd(crow + i, a);
                            else crow[i] = a;
                        }
                    }
                }
            }
        }
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem, BN);
}

// K-major bf16 matrix [rows, K] with row stride ld (elements) -> 2-D map, box 64(K) x 128(rows)
static inline int make_operand_map(CUtensorMap* m, const void* base, uint64_t rows, uint64_t K, uint64_t ld) {
    const uint64_t dims[2] = {K, rows};
    const uint64_t strides[1] = {ld * 2};
    const uint32_t box[2] = {(uint32_t)BK, 128u};
    return make_tmap_bf16(m, base, 2, dims, strides, box);
}

// output tile map: rows M (stride ldc elements), 128B-swizzled boxes of 128 rows x 128 bytes
static inline int make_output_map(CUtensorMap* m, void* base, int mode, uint64_t M, uint64_t N, uint64_t ldc) {
    const bool bf = mode == OUT_BF16;
    const uint64_t dims[2] = {N, M};
    const uint64_t strides[1] = {ldc * (bf ? 2u : 4u)};
    const uint32_t box[2] = {bf ? 64u : 32u, 128u};
    return make_tmap_typed(m, bf ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, 2, dims, strides, box);
}

// MN-major bf16 operand stored [K rows][MN] with row stride ld (elements) -> 2-D map, box 64(MN) x 64(K rows)
static inline int make_operand_map_mn(CUtensorMap* m, const void* base, uint64_t Krows, uint64_t MN, uint64_t ld) {
    const uint64_t dims[2] = {MN, Krows};
    const uint64_t strides[1] = {ld * 2};
    const uint32_t box[2] = {64u, (uint32_t)BK};
    return make_tmap_bf16(m, base, 2, dims, strides, box);
}

static inline cudaError_t launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const Params& p_in, cudaStream_t st,
                                 const CUtensorMap* tmAlo = nullptr, const CUtensorMap* tmBlo = nullptr) {
    {   // the shared-memory opt-in is per device: cache it per device ordinal, not process-wide
        static bool attr_set[64] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= 64 || !attr_set[dev]) {
            cudaError_t e = cudaFuncSetAttribute(gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes_for(MAX_STAGES));
            if (e != cudaSuccess) return e;
            if (dev >= 0 && dev < 64) attr_set[dev] = true;
        }
    }
    Params p = p_in;
    if (p.nsplit != 3 || !tmAlo || !tmBlo) p.nsplit = 1;
    if (p.blk.U == 0) { p.blk.U = 128; p.blk.NB = 16; }
    {   // ring depth from the k-blocks one CTA walks: short reductions are latency-bound per tile, so trade ring depth
        // for more co-resident CTAs (2 stages -> 3 CTAs/SM); long ones keep 3 stages (2 CTAs/SM)
        const int kblocks = (((p.K + BK - 1) / BK + p.splitk - 1) / p.splitk) * p.nsplit;
        p.stages = kblocks <= 2 ? kblocks : (kblocks <= 12 ? 2 : 3);
        if (p.stages < 1) p.stages = 1;
    }
    CUtensorMap tmC[2];
    memset(tmC, 0, sizeof(tmC));
    // TMA epilogue needs 16-byte aligned rows and base; otherwise the direct-store epilogue is used
    const size_t es = p.mode == OUT_BF16 ? 2 : 4;
    const bool scan_mode = p.mode >= OUT_SCAN_BF16;
    bool tma_ok = !scan_mode && p.batch <= 2 && ((p.ldc * es) % 16 == 0) && ((p.zC * es) % 16 == 0) && ((uintptr_t)p.C % 16 == 0);
    if (tma_ok) {
        for (int z = 0; z < p.batch; ++z)
            if (make_output_map(&tmC[z], (uint8_t*)p.C + (size_t)z * p.zC * es, p.mode, (uint64_t)p.M, (uint64_t)p.N, (uint64_t)p.ldc) != 0) tma_ok = false;
        if (p.batch == 1) tmC[1] = tmC[0];
    }
    p.tma_store = tma_ok ? 1 : 0;
    if (scan_mode && p.M % BM == 0 && p.blk.H % p.blk.U == 0 && p.blk.B % p.blk.NB == 0 && ((uintptr_t)p.C % 16 == 0)) p.tma_store = 1;   // bulk-store epilogue
    dim3 grid((p.N + BN - 1) / BN, (p.M + BM - 1) / BM, p.batch * p.splitk);
    if (p.m_fast) { const unsigned t = grid.x; grid.x = grid.y; grid.y = t; }
    const int stage_out = p.tma_store ? ((p.mode == OUT_BF16 || p.mode == OUT_SCAN_BF16) ? 32768 : 65536) : 0;
    gemm_kernel<<<grid, THREADS, smem_bytes_for(p.stages, stage_out), st>>>(tmA, tmB, p.nsplit == 3 ? *tmAlo : tmA, p.nsplit == 3 ? *tmBlo : tmB,
                                                                            tmC[0], tmC[1], p);
    return cudaGetLastError();
}

}  // namespace tcg
