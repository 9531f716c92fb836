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
#include "../../financial_market_data_analysis_b200/csrc/tc_common.cuh"
#include <cstring>

namespace tcg {

constexpr int BM = 128, BK = 64, MAX_STAGES = 6;             // per CTA: 128 rows of A and 128 rows of B per stage (32 KB)
constexpr int A_BYTES = BM * BK * 2, B_BYTES = 128 * BK * 2;
// Two tile configurations of the same kernel (template <CG, TBN>):
//   <1, 128>  one CTA, 128 x 128 tile, tcgen05.mma.cta_group::1 (M=128, N=128); 3-4 CTAs per SM
//   <2, 256>  a CTA PAIR (cluster 2x1x1) computes a 256 x 256 tile with tcgen05.mma.cta_group::2 (M=256, N=256): each CTA
//             loads its own 128 rows of A and HALF of the B tile (128 rows) - the same 32 KB per stage as <1,128> for four
//             times the MMA work of a k-block per pair, which halves the operand bytes an SM has to ingest per FLOP (the
//             128 x 128 tile is bound by exactly that, ~64 B/clk/SM from L2).  The leader CTA (rank 0) issues the MMAs for
//             both; each CTA keeps its 128 x 256 accumulator half in its own tensor memory and runs its own epilogue.
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
struct ScanBlk { int T, B, H, G; };

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
    int persist;              // in: 1 = ask for the persistent form; launch() resolves it to 0/1 (see gemm_kernel)
    int tiles_m, tiles_n, work;   // filled by launch(): tile grid and number of work items
    int pair;                 // 1: 256 x 256 tiles on CTA pairs (tcgen05 cta_group::2); 0: 128 x 128 single-CTA tiles
    unsigned int* dbg;        // watchdog record (nullable)
};

template <int CG, int TBN>
__device__ __forceinline__ void tile_coords(const Params& p, int w, uint32_t rank, int& m0, int& n0, int& z, int& ks, int& kb0, int& nkb) {
    int mt, nt, zz;
    if (p.m_fast || CG == 2) { mt = w % p.tiles_m; const int r = w / p.tiles_m; nt = r % p.tiles_n; zz = r / p.tiles_n; }
    else { nt = w % p.tiles_n; const int r = w / p.tiles_n; mt = r % p.tiles_m; zz = r / p.tiles_m; }
    z = zz / p.splitk; ks = zz % p.splitk;
    m0 = mt * BM * CG + (int)rank * BM; n0 = nt * TBN;
    const int kblocks_total = (p.K + BK - 1) / BK;
    const int kb_per = (kblocks_total + p.splitk - 1) / p.splitk;
    kb0 = ks * kb_per;
    nkb = max(0, min(kblocks_total, kb0 + kb_per) - kb0);
}

// Work items (output tiles x batch x split-K slices) are walked by a 1-D grid: CTA (pair) c takes items c, c + G, ...
// With p.persist the grid is one CTA per SM, the smem ring runs ahead across tiles, the accumulator is double-buffered
// in tensor memory and the epilogue (own staging buffer) of tile i overlaps the main loop of tile i + 1; without it
// every CTA has exactly one item and the staging buffer aliases the (then idle) ring.
template <int CG, int TBN>
__global__ void __launch_bounds__(THREADS, CG == 1 ? 4 : 2)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmC0, const __grid_constant__ CUtensorMap tmC1, const Params p) {
    constexpr int BN = 128;                    // column block of the epilogue (TBN / 128 blocks per tile)
    constexpr int NCB = TBN / BN;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int STAGES = p.stages;
    // the TMA-store epilogue needs up to 64 KB of staging (fp32 tile)
    const int ring_bytes = STAGES * (A_BYTES + B_BYTES);
    const int stage_out = p.tma_store ? ((p.mode == OUT_BF16 || p.mode == OUT_SCAN_BF16) ? 32768 : 65536) : 0;
    const int data_bytes = p.persist ? ring_bytes + stage_out : (ring_bytes < stage_out ? stage_out : ring_bytes);
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * A_BYTES;
    uint8_t* stg = p.persist ? smem + ring_bytes : smem;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + data_bytes);
    uint64_t* full = bars;                 // [STAGES] TMA -> MMA
    uint64_t* empty = bars + 8;            // [STAGES] MMA -> TMA
    uint64_t* tfull = bars + 16;           // [2] MMA -> epilogue (accumulator stage complete)
    uint64_t* tempty = bars + 18;          // [2] epilogue -> MMA (accumulator stage drained), one arrival per epilogue warp
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
    const uint32_t tmem_cols = p.persist ? 2u * TBN : (uint32_t)TBN;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = CG == 2 ? tc::cluster_ctarank() : 0u;      // CG == 2: consecutive CTAs form the pair
    const int w0 = CG == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int wstride = CG == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { tc::mbar_init(&full[s], CG); tc::mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { tc::mbar_init(&tfull[a], 1); tc::mbar_init(&tempty[a], 4 * CG); }
        tc::fence_mbar_init();
    }
    if (warp == 1) { if (CG == 2) tc::tmem_alloc_cg2(tmem_slot, tmem_cols); else tc::tmem_alloc(tmem_slot, tmem_cols); }
    tc::tcgen05_fence_before();
    __syncthreads();
    if (CG == 2) tc::cluster_sync_all();       // the peer's barriers exist before anything signals them
    tc::tcgen05_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (warp == 0 && lane == 0) { tc::tma_prefetch_desc(&tmA); tc::tma_prefetch_desc(&tmB); }

    if (warp == 0) {
        if (tc::elect_one()) {
            tc::tma_prefetch_desc(&tmA);
            tc::tma_prefetch_desc(&tmB);
            int g = 0;                             // k-blocks issued so far: the ring runs on across tiles
            bool ok = true;
            for (int w = w0; w < p.work && ok; w += wstride) {
                int m0, n0, z, ks, kb0, nkb;
                tile_coords<CG, TBN>(p, w, rank, m0, n0, z, ks, kb0, nkb);
                const int nB0 = n0 + (int)rank * 128;      // first B row this CTA loads (its half of the pair's B tile)
                for (int i = 0; i < nkb; ++i, ++g) {
                    const int s = g % STAGES;
                    const uint32_t ph = (g / STAGES) & 1;
                    if (!tc::mbar_wait(&empty[s], ph ^ 1, p.dbg, 0x100 + s)) { ok = false; break; }
                    const int k = (kb0 + i) * BK;
                    if (CG == 2) {
                        // both CTAs load into their own smem; every byte is counted on the LEADER's full barrier
                        const uint32_t lbar = tc::mapa_u32(tc::smem_u32(&full[s]), 0);
                        if (rank == 0) tc::mbar_arrive_expect_tx(&full[s], 2 * (A_BYTES + B_BYTES));
                        else tc::mbar_arrive_cluster(&full[s], 0);
                        if (p.a_mn) {
                            tc::tma_load_2d_cg2(sA + s * A_BYTES, &tmA, lbar, p.a_row_off[z] + m0, k);
                            tc::tma_load_2d_cg2(sA + s * A_BYTES + A_BYTES / 2, &tmA, lbar, p.a_row_off[z] + m0 + 64, k);
                        } else tc::tma_load_2d_cg2(sA + s * A_BYTES, &tmA, lbar, k, p.a_row_off[z] + m0);
                        if (p.b_mn) {
                            tc::tma_load_2d_cg2(sB + s * B_BYTES, &tmB, lbar, p.b_row_off[z] + nB0, k + p.b_k_off[z]);
                            tc::tma_load_2d_cg2(sB + s * B_BYTES + B_BYTES / 2, &tmB, lbar, p.b_row_off[z] + nB0 + 64, k + p.b_k_off[z]);
                        } else tc::tma_load_2d_cg2(sB + s * B_BYTES, &tmB, lbar, k + p.b_k_off[z], p.b_row_off[z] + nB0);
                        continue;
                    }
                    tc::mbar_arrive_expect_tx(&full[s], A_BYTES + B_BYTES);
                    if (p.a_mn) {        // two boxes of 64 (MN) x 64 (K rows)
                        tc::tma_load_2d(sA + s * A_BYTES, &tmA, &full[s], p.a_row_off[z] + m0, k);
                        tc::tma_load_2d(sA + s * A_BYTES + A_BYTES / 2, &tmA, &full[s], p.a_row_off[z] + m0 + 64, k);
                    } else tc::tma_load_2d(sA + s * A_BYTES, &tmA, &full[s], k, p.a_row_off[z] + m0);
                    if (p.b_mn) {
                        tc::tma_load_2d(sB + s * B_BYTES, &tmB, &full[s], p.b_row_off[z] + n0, k + p.b_k_off[z]);
                        tc::tma_load_2d(sB + s * B_BYTES + B_BYTES / 2, &tmB, &full[s], p.b_row_off[z] + n0 + 64, k + p.b_k_off[z]);
                    } else tc::tma_load_2d(sB + s * B_BYTES, &tmB, &full[s], k + p.b_k_off[z], p.b_row_off[z] + n0);
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0 && tc::elect_one()) {       // CG == 2: the leader issues for the pair
            const uint32_t idesc = tc::umma_idesc_bf16(BM * CG, TBN, (uint32_t)p.a_mn, (uint32_t)p.b_mn);
            int g = 0, it = 0;
            bool ok = true;
            for (int w = w0; w < p.work && ok; w += wstride, ++it) {
                int m0, n0, z, ks, kb0, nkb;
                tile_coords<CG, TBN>(p, w, rank, m0, n0, z, ks, kb0, nkb);
                const int a = p.persist ? (it & 1) : 0;
                if (!tc::mbar_wait(&tempty[a], ((it >> 1) & 1) ^ 1, p.dbg, 0x280 + a)) break;     // epilogue drained this stage
                tc::tcgen05_fence_after();
                const uint32_t tacc = tmem + (uint32_t)(a * TBN);
                for (int i = 0; i < nkb; ++i, ++g) {
                    const int s = g % STAGES;
                    const uint32_t ph = (g / STAGES) & 1;
                    if (!tc::mbar_wait(&full[s], ph, p.dbg, 0x200 + s)) { ok = false; break; }
                    tc::tcgen05_fence_after();
                    const uint32_t aa = tc::smem_u32(sA + s * A_BYTES), ab = tc::smem_u32(sB + s * B_BYTES);
                    const uint64_t da = p.a_mn ? tc::umma_desc_mn_sw128(aa, A_BYTES / 2) : tc::umma_desc_k_sw128(aa);
                    const uint64_t db = p.b_mn ? tc::umma_desc_mn_sw128(ab, B_BYTES / 2) : tc::umma_desc_k_sw128(ab);
                    // per K=16 step: K-major advances 32 B inside the swizzled row; MN-major advances 16 K-rows = 2048 B
                    const uint64_t sa = p.a_mn ? 128 : 2, sb = p.b_mn ? 128 : 2;
#pragma unroll
                    for (int kk = 0; kk < BK / 16; ++kk) {
                        if (CG == 2) tc::umma_bf16_cg2(tacc, da + sa * kk, db + sb * kk, idesc, (i | kk) ? 1u : 0u);
                        else tc::umma_bf16(tacc, da + sa * kk, db + sb * kk, idesc, (i | kk) ? 1u : 0u);
                    }
                    if (CG == 2) tc::umma_commit_mc2(&empty[s]); else tc::umma_commit(&empty[s]);   // frees the smem slot(s)
                }
                if (CG == 2) tc::umma_commit_mc2(&tfull[a]); else tc::umma_commit(&tfull[a]);
            }
        }
    } else {
        // epilogue warps 2..5 -> TMEM lane quarter = warp % 4
        const int q = warp & 3;
        int it = 0;
        for (int w = w0; w < p.work; w += wstride, ++it) {
            int m0, n0, z, ks, kb0, nkb;
            tile_coords<CG, TBN>(p, w, rank, m0, n0, z, ks, kb0, nkb);
            const int a = p.persist ? (it & 1) : 0;
            const uint32_t tacc = tmem + (uint32_t)(a * TBN);
            const int m = m0 + q * 32 + lane;
            bool ok = tc::mbar_wait(&tfull[a], (it >> 1) & 1, p.dbg, 0x300);
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
                for (int cb = 0; cb < NCB; ++cb) {
                    const int nc0 = n0 + cb * BN;                 // first column of this 128-column block
                    if (nc0 >= p.N) break;
                    if (cb > 0 || it > 0) asm volatile("bar.sync 1, 128;" ::: "memory");     // the previous block's smem has been read
    #pragma unroll 1
                    for (int c = 0; c < BN / 32; ++c) {
                        uint32_t v[32];
                        if (nkb > 0) { tc::tmem_ld32(tacc + ((uint32_t)(q * 32) << 16) + cb * BN + c * 32, v); tc::tmem_ld_wait(); }
                        else {
    #pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] = 0u;
                        }
                        const int nb = nc0 + c * 32;
    #pragma unroll
                        for (int i8 = 0; i8 < 32; i8 += 8) {
                            float f[8];
    #pragma unroll
                            for (int j = 0; j < 8; ++j)
                                f[j] = __uint_as_float(v[i8 + j]) + (bias ? (p.bias_per_row ? brow : ((nb + i8 + j < p.N) ? bias[nb + i8 + j] : 0.f)) : 0.f);
                            uint8_t* dst = stg + (size_t)(c * 4 + (i8 >> 3)) * run_bytes;
                            if (bf) {
                                uint32_t w[4];
    #pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
                                    w[j] = *reinterpret_cast<uint32_t*>(&h2);
                                }
                                *reinterpret_cast<uint4*>(dst + ml * 16) = make_uint4(w[0], w[1], w[2], w[3]);
                            } else {
                                *reinterpret_cast<float4*>(dst + ml * 32) = make_float4(f[0], f[1], f[2], f[3]);
                                *reinterpret_cast<float4*>(dst + ml * 32 + 16) = make_float4(f[4], f[5], f[6], f[7]);
                            }
                        }
                    }
                    tc::fence_proxy_async_smem();
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    if (warp == 2 && tc::elect_one()) {
                        const int H = p.blk.H, G = p.blk.G, Bb = p.blk.B, Tt = p.blk.T;
                        const int dd = m0 / (G * H), gg = (m0 / H) % G, cc = (m0 % H) / 128;
                        const int CSs = H / 128, ntl = Bb / 16;
                        const size_t es = bf ? 2 : 4;
                        for (int r = 0; r < BN / 8; ++r) {
                            const int n = nc0 + r * 8;
                            if (n >= p.N) break;
                            const int t_ = n / Bb, b = n % Bb;
                            const int tile_ = b >> 4, half = (b >> 3) & 1;
                            const size_t e = ((((((size_t)dd * ntl + tile_) * Tt + t_) * CSs + cc) * G + gg) * 256 + (size_t)half * 128) * 8;
                            tc::bulk_s2g(reinterpret_cast<uint8_t*>(p.C) + e * es, stg + (size_t)r * run_bytes, (uint32_t)run_bytes);
                        }
                        tc::tma_store_commit();
                        tc::tma_store_wait_read();     // smem may be released once it has been read; the writes complete on their own
                    }
                }
            } else if (ok && p.tma_store) {
                // ---- staged epilogue: TMEM -> registers -> 128B-swizzled smem boxes -> TMA store / reduce-add.
                // The pipeline stages are free (every MMA has retired), so they serve as the staging buffer:
                // bf16: 2 boxes of [128 rows x 64 cols], fp32: 4 boxes of [128 rows x 32 cols], 16 KB each.
                const int ml = q * 32 + lane;
                const uint32_t sw = (uint32_t)(ml & 7);
                const bool is_bf16 = p.mode == OUT_BF16;
    #pragma unroll 1
                for (int cb = 0; cb < NCB; ++cb) {
                    const int nc0 = n0 + cb * BN;
                    if (nc0 >= p.N) break;
                    if (cb > 0 || it > 0) asm volatile("bar.sync 1, 128;" ::: "memory");
    #pragma unroll 1
                    for (int c = 0; c < BN / 32; ++c) {
                        uint32_t v[32];
                        if (nkb > 0) {
                            tc::tmem_ld32(tacc + ((uint32_t)(q * 32) << 16) + cb * BN + c * 32, v);
                            tc::tmem_ld_wait();
                        } else {
    #pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] = 0u;
                        }
                        const int nb = nc0 + c * 32;
                        if (bias && ks == 0) {
    #pragma unroll
                            for (int i = 0; i < 32; ++i) {
                                const float bb = p.bias_per_row ? brow : ((nb + i < p.N) ? bias[nb + i] : 0.f);
                                v[i] = __float_as_uint(__uint_as_float(v[i]) + bb);
                            }
                        }
                        if (is_bf16) {
                            uint8_t* box = stg + (size_t)(c >> 1) * 16384 + (size_t)ml * 128;
    #pragma unroll
                            for (int k4 = 0; k4 < 4; ++k4) {
                                uint32_t w[4];
    #pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(v[k4 * 8 + 2 * j]), __uint_as_float(v[k4 * 8 + 2 * j + 1]));
                                    w[j] = *reinterpret_cast<uint32_t*>(&h2);
                                }
                                const uint32_t chunk = ((uint32_t)((c & 1) * 4 + k4)) ^ sw;
                                *reinterpret_cast<uint4*>(box + chunk * 16) = make_uint4(w[0], w[1], w[2], w[3]);
                            }
                        } else {
                            uint8_t* box = stg + (size_t)c * 16384 + (size_t)ml * 128;
    #pragma unroll
                            for (int k4 = 0; k4 < 8; ++k4) {
                                const uint32_t chunk = ((uint32_t)k4) ^ sw;
                                *reinterpret_cast<uint4*>(box + chunk * 16) = make_uint4(v[k4 * 4], v[k4 * 4 + 1], v[k4 * 4 + 2], v[k4 * 4 + 3]);
                            }
                        }
                    }
                    tc::fence_proxy_async_smem();
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    if (warp == 2 && tc::elect_one()) {
                        const CUtensorMap* tmC = z == 0 ? &tmC0 : &tmC1;
                        const int nboxes = is_bf16 ? BN / 64 : BN / 32;
                        const int bw = is_bf16 ? 64 : 32;
                        for (int b = 0; b < nboxes; ++b) {
                            if (nc0 + b * bw >= p.N) break;
                            if (p.mode == OUT_ATOMIC_F32) tc::tma_reduce_add_2d(tmC, stg + (size_t)b * 16384, nc0 + b * bw, m0);
                            else tc::tma_store_2d(tmC, stg + (size_t)b * 16384, nc0 + b * bw, m0);
                        }
                        tc::tma_store_commit();
                        tc::tma_store_wait_read();     // smem may be released once it has been read; the writes complete on their own
                    }
                }
            } else if (ok) {
    #pragma unroll 1
                for (int c = 0; c < TBN / 32; ++c) {
                    uint32_t v[32];
                    if (nkb > 0) {
                        tc::tmem_ld32(tacc + ((uint32_t)(q * 32) << 16) + c * 32, v);
                        tc::tmem_ld_wait();
                    } else {
    #pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = 0u;
                    }
                    const int nb = n0 + c * 32;
                    if (m < p.M && nb < p.N && p.mode >= OUT_SCAN_BF16) {
                        // blocked store: 8 consecutive b (one thread-run of the scan kernel) = one 16/32-byte store
                        const int H = p.blk.H, G = p.blk.G, Bb = p.blk.B, Tt = p.blk.T;
                        const int dd = m / (G * H), gg = (m / H) % G, unit = m % H;
                        const int CSs = H / 128, cc = unit / 128, ju = unit % 128;
                        const int ntl = Bb / 16;
    #pragma unroll
                        for (int i8 = 0; i8 < 32; i8 += 8) {
                            const int n = nb + i8;
                            if (n >= p.N) break;
                            const int t = n / Bb, b = n % Bb;
                            const int tile = b >> 4, half = (b >> 3) & 1;
                            const int tid = ((ju >> 5) + 4 * half) * 32 + (ju & 31);
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
            // this accumulator stage has been read: the MMA warp may overwrite it (tile it + 2)
            tc::tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) { if (CG == 2) tc::mbar_arrive_cluster(&tempty[a], 0); else tc::mbar_arrive(&tempty[a]); }
            if (!ok) break;
        }
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    if (CG == 2) tc::cluster_sync_all();       // the leader's MMAs read the peer's smem: nobody leaves early
    if (warp == 1) { if (CG == 2) tc::tmem_dealloc_cg2(tmem, tmem_cols); else tc::tmem_dealloc(tmem, tmem_cols); }
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

static inline cudaError_t launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const Params& p_in, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_kernel<1, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_kernel<2, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    Params p = p_in;
    const int kblocks = ((p.K + BK - 1) / BK + p.splitk - 1) / p.splitk;
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
    if (scan_mode && p.M % BM == 0 && p.blk.B % 16 == 0 && ((uintptr_t)p.C % 16 == 0)) p.tma_store = 1;   // bulk-store epilogue
    const int stage_out = p.tma_store ? ((p.mode == OUT_BF16 || p.mode == OUT_SCAN_BF16) ? 32768 : 65536) : 0;
    const int cg = p.pair ? 2 : 1, tbn = p.pair ? 256 : 128;
    p.tiles_m = (p.M + cg * BM - 1) / (cg * BM);
    p.tiles_n = (p.N + tbn - 1) / tbn;
    p.work = p.tiles_m * p.tiles_n * p.batch * p.splitk;
    static int n_sm = 0;
    if (!n_sm) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); if (n_sm <= 0) n_sm = 148; }
    // persistent form when there are several tiles per SM (pair tiles: per SM pair): deep ring + separate staging + two
    // accumulator stages; otherwise one tile per CTA with a shallow ring so that several CTAs share an SM
    const int slots = p.pair ? n_sm / 2 : n_sm;
    const int per_sm = p.persist > 1 ? p.persist : 1;
    p.persist = (p.persist > 0 && p.work >= 2 * slots) ? 1 : 0;
    unsigned ctas;
    if (p.persist) {
        const int avail = ((227 * 1024) / per_sm - 1024 - 256 - 1024 - stage_out) / (A_BYTES + B_BYTES);
        p.stages = avail > MAX_STAGES ? MAX_STAGES : avail;
        if (p.stages < 1) p.stages = 1;
        ctas = (unsigned)(p.work < slots * per_sm ? p.work : slots * per_sm) * cg;
    } else {
        if (p.pair) p.stages = kblocks <= 2 ? (kblocks < 1 ? 1 : kblocks) : 3;          // 2 CTAs per SM (2 x 256 TMEM columns)
        else {
            // ring depth from the k-blocks one CTA walks: short reductions are latency-bound per tile, so trade ring depth
            // for more co-resident CTAs (2 stages -> 3 CTAs/SM); long ones keep 3 stages (2 CTAs/SM)
            p.stages = kblocks <= 2 ? kblocks : (kblocks <= 12 ? 2 : 3);
            if (p.stages < 1) p.stages = 1;
        }
        ctas = (unsigned)p.work * cg;
    }
    const size_t smem = p.persist ? (size_t)p.stages * (A_BYTES + B_BYTES) + stage_out + 1024 + 256 : (size_t)smem_bytes_for(p.stages, stage_out);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(ctas);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cg; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (p.pair) return cudaLaunchKernelEx(&cfg, gemm_kernel<2, 256>, tmA, tmB, tmC[0], tmC[1], p);
    return cudaLaunchKernelEx(&cfg, gemm_kernel<1, 128>, tmA, tmB, tmC[0], tmC[1], p);
}

}  // namespace tcg
