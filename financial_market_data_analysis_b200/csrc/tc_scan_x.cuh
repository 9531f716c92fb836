// tc_scan_x.cuh - fp32-class persistent GRU recurrence on tcgen05 tensor cores ("x3": every operand is split into a
// bf16 high part and a bf16 low part, x = hi + lo with |x - hi - lo| <= 2^-17 |x|, products accumulate in fp32).
// This is the variant that meets the reference's fp32 tolerance (logits <= 1e-4 rel of torch.nn.GRU,
// /root/reference/biGRU_model.py:54-56,102): bf16 operands alone stop at ~3e-3 through 128-256 recurrent steps.
//
// Tensor memory holds 256 KB per SM, i.e. 64 hidden units x 3 gates x H=256 weights as (hi, lo) bf16 pairs.  So one
// thread-block CLUSTER of CS = H/64 CTAs walks all T steps of one (direction, 32-row batch tile); CTA c owns hidden units
// [64c, 64c+64).  Gate math, state and stash are fp32; exp/reciprocal use ex2.approx / rcp.approx (<= 2^-22 rel).
//
// Forward (output-partitioned, "all-gather"): the A operand rows are STACKED, rows 0-63 = hi(W_hh[unit]), rows 64-127 =
// lo(W_hh[unit]) (M = 128), the B operand is the batch tile [32 x H] of h_{t-1}, once as hi and once as lo:
//     D_g[row, b] = sum_k A_g[row, k] (h_hi + h_lo)[b, k]    ->   W h = D[unit] + D[64 + unit]        (4-term product)
// The two halves live in different TMEM lanes, so the epilogue warps of lanes 64-127 and 0-63 swap half of their columns
// through shared memory; every thread then owns one unit x 8 batch columns.  New h (hi, lo) goes into the operand tile of
// the next step, locally and into every peer CTA with st.async (one mbarrier per source CTA, so a peer's K chunk is
// multiplied as soon as it lands).
//
// Backward (reduction-partitioned, "reduce-scatter"): shipping dgh (3H wide, hi + lo) to every peer would move 3x the
// forward's bytes through DSMEM.  Instead CTA c keeps the W_hh rows of ITS OWN units' gates (q in own 3 x 64) for ALL H
// output units k as the A operand - row blocks (part p in {hi, lo}) x (k half) of 128 lanes x K = 192 - multiplies them with
// its local dgh tile [32 x 192] (hi, then lo) and sends the fp32 partial sums of units it does not own to their owners
// (st.async into a receive buffer); hi/lo row blocks of one k share a lane, so their sum is formed in registers.
//
// Kernels in this file: gru_scanx_fwd_kernel / gru_scanx_bwd_kernel (single 32-row tile per cluster: H = 128, and H = 256 on
// request) and their PING-PONG forms gru_scanx_fwd2_kernel / gru_scanx_bwd2_kernel (H = 256, the default): the tile is worked as
// two 16-row sub-tiles that alternate on the tensor pipe, the epilogue warps and the DSMEM network (see the comments there).
//
// Blocked ("scan-private") layouts (time-major, fp32): block (d, tile, t, cta) = (((d*ntiles + tile)*T + t)*CS + cta),
// inside a block [gate][thread 0..255][8 batch columns]; thread tid = j + 64*(cb/8) <-> unit j = tid % 64 of the CTA,
// batch columns [8*(tid/64), +8) of the 32-row tile.
//   giX  fp32 [block][3][256][8]   input projection incl. b_ih (+ b_hh for r, z)           (tc_gemm OUT_SCAN_F32)  read
//   GX   fp32 [block][4][256][8]   r, z, n, W_hn h + b_hn                                   stash, written fwd / read bwd
//   YBX  fp32 [block][256][8]      h_t                                                      written fwd / read bwd
//   dYBX fp32 [block][256][8]      dL/dy_t of this layer (lower layers)                     (tc_gemm OUT_SCAN_F32)  read
//   Yhi / Ylo bf16 [R][D*H]        layer output split, row-major (next layer's GEMM operands, head)         written
//   dgi_hi/lo bf16 [R][D*3H], dghn_hi/lo bf16 [R][D*H]                                     written bwd (GEMM operands)
#pragma once
#include "tc_common.cuh"
#include "tc_scan.cuh"

namespace tcx {

#ifdef BIGRU_SCAN_TIMING
#define SCANX_TS(slot) do { if (blockIdx.x == 0 && s >= 64 && s < 72) p.ts[(s - 64) * 16 + (slot)] = clock64(); } while (0)
#else
#define SCANX_TS(slot) do { } while (0)
#endif

constexpr int NB = 32;            // batch rows per tile = UMMA N
constexpr int UNITS = 64;         // hidden units per CTA
constexpr int EPI_WARPS = 8;
constexpr int THREADS = (EPI_WARPS + 2) * 32;
constexpr int H_CHUNK = NB * 128;          // bytes of one [32 x 64] bf16 K-major chunk (128B swizzle)
constexpr int GI_BLOCK = 3 * 256 * 32;     // fp32 blocks
constexpr int G_BLOCK = 4 * 256 * 32;
constexpr int YB_BLOCK = 256 * 32;
constexpr int DY_BLOCK = 256 * 32;
constexpr int XBUF_BYTES = 8 * 6 * 32 * 16;   // forward lane-half exchange: [warp][gate*2+k][lane] float4
constexpr int NSF = 3, NSB = 2;
constexpr uint32_t A_COL = 128;            // accumulators in columns [0, 128), weights from column 128 (384 columns at H=256)
constexpr int BWD_STAGE = G_BLOCK + YB_BLOCK + DY_BLOCK;
constexpr int RECV_BYTES = 4 * 8 * 64 * 16;   // backward partial sums: [src cta][column group of 4][unit j] float4

__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16(x);
    lo = __float2bfloat16(x - __bfloat162float(hi));
}

__device__ __forceinline__ void tmem_ld16f(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    tc::tmem_ld16(taddr, r);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// the registers of a tcgen05.ld are defined only after tcgen05.wait::ld: tie them to the wait so that no consumer (not even a
// register move) can be scheduled between the load and the wait
__device__ __forceinline__ void tmem_ld_wait_pin(float (&a)[16], float (&b)[16], float (&c)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) asm volatile("" : "+f"(a[i]), "+f"(b[i]), "+f"(c[i]));
}
__device__ __forceinline__ void tmem_ld_wait_pin(float (&a)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) asm volatile("" : "+f"(a[i]));
}
__device__ __forceinline__ void tmem_ld_wait_pin(float (&a)[16], float (&b)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) asm volatile("" : "+f"(a[i]), "+f"(b[i]));
}

__device__ __forceinline__ size_t blk_index(int d, int tile, int t, int c, int ntiles, int T, int CS) {
    return (((size_t)d * ntiles + tile) * T + t) * CS + c;
}
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory"); }

static inline size_t fwd_smem_bytes(int H) {
    const int KC = H / 64;
    return (size_t)2 * 2 * KC * H_CHUNK + (size_t)NSF * GI_BLOCK + XBUF_BYTES + 1024 + 512;
}
static inline size_t bwd_smem_bytes() {
    return (size_t)2 * 2 * 3 * H_CHUNK + (size_t)2 * 2 * H_CHUNK + (size_t)2 * RECV_BYTES + (size_t)NSB * BWD_STAGE + 1024 + 512;
}

struct FwdParams {
    int B, T, H, D;
    const __nv_bfloat16* Wimg;    // [D][CS][128 rows: hi units, lo units][3][H]
    const float* giX;
    const float* b_hn;            // [D][H]
    const float* h0;              // nullable [D][B][H]: initial hidden state of this layer ...
    const float* gh0;             // ... and its recurrent product W_hh h0 [D][B][3H] (fp32, formed by the caller): step 0 reads it
                                  // instead of a tensor-core product, so the scan itself always starts from step 1
    float* GX;
    float* YBX;
    float* hn_out;                // nullable [D][B][H]
    __nv_bfloat16* Yhi;           // [R][D*H]
    __nv_bfloat16* Ylo;
    unsigned int* dbg;
    CUtensorMap tmYhi, tmYlo;     // box 64 x 32 (filled by launch_fwd)
#ifdef BIGRU_SCAN_TIMING
    unsigned long long* ts;
#endif
};

// K chunk `u` (64 columns of h, hi part then lo part) of all three gates: 24 MMAs, fully unrolled
template <int H, bool FIRST>
__device__ __forceinline__ void fwd_issue_chunk(uint32_t tmem_d, uint32_t tmem_a_chunk, uint64_t desc_hi, uint64_t desc_lo) {
    constexpr uint32_t idesc = tc::umma_idesc_bf16(128, NB);
#pragma unroll
    for (int g = 0; g < 3; ++g) {
#pragma unroll
        for (int part = 0; part < 2; ++part) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
                tcs::umma_bf16_ts(tmem_d + (uint32_t)(g * NB), tmem_a_chunk + (uint32_t)(g * (H / 2) + kk * 8),
                                  (part ? desc_lo : desc_hi) + (uint64_t)(2 * kk), idesc, (FIRST && part == 0 && kk == 0) ? 0u : 1u);
        }
    }
}

template <int H>
__global__ void __launch_bounds__(THREADS, 1) gru_scanx_fwd_kernel(const __grid_constant__ FwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int KC = H / 64, CS = KC;
    constexpr int TILE_BYTES = KC * H_CHUNK;               // one part (hi or lo) of one h operand tile
    const int B = p.B, T = p.T;
    uint8_t* sH = smem;                                    // [2 buf][2 part][KC][H_CHUNK]
    uint8_t* sIn = sH + (size_t)4 * TILE_BYTES;            // [NSF][GI_BLOCK]
    uint8_t* sX = sIn + (size_t)NSF * GI_BLOCK;            // exchange buffer
    uint64_t* bars = reinterpret_cast<uint64_t*>(sX + XBUF_BYTES);
    uint64_t* h_full = bars;                 // [2 buf][4 src]
    uint64_t* mma_done = bars + 8;
    uint64_t* epi_done = bars + 9;
    uint64_t* in_full = bars + 10;           // [NSF]
    uint64_t* in_empty = bars + 10 + NSF;    // [NSF]
    uint64_t* xch = bars + 10 + 2 * NSF;     // [8] per epilogue warp: "my half of the lane-half swap is in shared memory"
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18 + 2 * NSF);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t c = CS > 1 ? tc::cluster_ctarank() : 0u;
    const int cluster_id = blockIdx.x / CS;
    const int ntiles = B / NB;
    const int d = cluster_id / ntiles, tile = cluster_id % ntiles;
    const bool has_h0 = p.h0 != nullptr && p.gh0 != nullptr;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) tc::mbar_init(&h_full[i], 1);
        tc::mbar_init(mma_done, 1);
        tc::mbar_init(epi_done, EPI_WARPS);
        for (int i = 0; i < NSF; ++i) { tc::mbar_init(&in_full[i], 1); tc::mbar_init(&in_empty[i], EPI_WARPS); }
        for (int i = 0; i < EPI_WARPS; ++i) tc::mbar_init(&xch[i], 1);
        // first use of the per-source "peer chunk landed" barriers (h_s lands in buffer s & 1): armed here, before the
        // cluster-wide sync below, so that a fast peer's st.async bytes can never reach a barrier that does not expect them
        if (CS > 1)
            for (uint32_t u = 0; u < (uint32_t)CS; ++u) {
                if (u == c) continue;
                if (T > 1) tc::mbar_arrive_expect_tx(&h_full[u], 2 * H_CHUNK);         // h_0
                if (T > 2) tc::mbar_arrive_expect_tx(&h_full[4 + u], 2 * H_CHUNK);     // h_1
            }
        tc::fence_mbar_init();
    }
    if (warp == EPI_WARPS) tc::tmem_alloc(tmem_slot, 512);
    tc::tcgen05_fence_before();
    __syncthreads();
    if (CS > 1) tc::cluster_sync_all();
    tc::tcgen05_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (warp < EPI_WARPS) tcs::load_weights_to_tmem(p.Wimg + ((size_t)d * CS + c) * 128 * 3 * H, 3 * H, tmem, A_COL, warp, lane);
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();

    if (warp == EPI_WARPS + 1) {
        // ---- input prefetch: one bulk copy (24 KB) per step into the ring
        if (tc::elect_one()) {
            bool ok = true;
            for (int s = 0; s < T; ++s) {
                const int st = s % NSF;
                if (s >= NSF && ok) ok = tc::mbar_wait(&in_empty[st], ((s / NSF) - 1) & 1, p.dbg, 0x1300 + (s & 0xff));
                const int t = d == 0 ? s : T - 1 - s;
                tc::mbar_arrive_expect_tx(&in_full[st], GI_BLOCK);
                tc::bulk_g2s(sIn + (size_t)st * GI_BLOCK,
                             reinterpret_cast<const uint8_t*>(p.giX) + blk_index(d, tile, t, (int)c, ntiles, T, CS) * GI_BLOCK,
                             GI_BLOCK, &in_full[st]);
            }
        }
    } else if (warp == EPI_WARPS) {
        // ---- control thread: issues the 24*KC MMAs of a step; K chunk u is ready when source CTA u's bytes have landed
        if (tc::elect_one()) {
            bool ok = true;
            // h_s lands in buffer s & 1 and is multiplied at step s + 1 (step 0 has no product: h_-1 = 0, or the caller supplies
            // W_hh h0).  The first use of every barrier was armed at initialisation, every later one right after the previous
            // use was consumed; uses are counted, so the phase parities need no case analysis.
            uint32_t epi_rounds = 0, hf_use0 = 0, hf_use1 = 0;
            auto store_tile = [&](int step) {             // this CTA's 64 columns of Y (hi, lo) for time step `step`
                const int tt = d == 0 ? step : T - 1 - step;
                const uint8_t* src = sH + (size_t)(step & 1) * 2 * TILE_BYTES + (size_t)c * H_CHUNK;
                tc::tma_store_2d(&p.tmYhi, src, d * H + (int)c * UNITS, tt * B + tile * NB);
                tc::tma_store_2d(&p.tmYlo, src + TILE_BYTES, d * H + (int)c * UNITS, tt * B + tile * NB);
                tc::tma_store_commit();
            };
            const uint32_t hb0 = tc::smem_u32(sH);
            for (int s = 1; s < T; ++s) {
                const int pb = (s - 1) & 1;
                const uint32_t tb = hb0 + (uint32_t)pb * 2 * TILE_BYTES;
                if (ok) ok = tc::mbar_wait(epi_done, epi_rounds & 1, p.dbg, 0x1400 + (s & 0xff));
                ++epi_rounds;
                SCANX_TS(0);
                tc::tcgen05_fence_after();
                // own chunk first (it is local), then the peers' chunks in ring order as they land
                fwd_issue_chunk<H, true>(tmem, tmem + A_COL + c * 32, tc::umma_desc_k_sw128(tb + c * H_CHUNK),
                                         tc::umma_desc_k_sw128(tb + TILE_BYTES + c * H_CHUNK));
                SCANX_TS(1);
                for (uint32_t i = 1; i < (uint32_t)CS; ++i) {
                    const uint32_t u = (c + i) % CS;
                    if (ok) ok = tc::mbar_wait(&h_full[pb * 4 + u], (pb ? hf_use1 : hf_use0) & 1, p.dbg, 0x1500 + (s & 0xff));
                    if (s + 2 < T) tc::mbar_arrive_expect_tx(&h_full[pb * 4 + u], 2 * H_CHUNK);      // h_{s+1} comes to this buffer
                    tc::tcgen05_fence_after();
                    fwd_issue_chunk<H, false>(tmem, tmem + A_COL + u * 32, tc::umma_desc_k_sw128(tb + u * H_CHUNK),
                                              tc::umma_desc_k_sw128(tb + TILE_BYTES + u * H_CHUNK));
                }
                if (pb) ++hf_use1; else ++hf_use0;
                tc::tma_store_wait_read();
                tc::umma_commit(mma_done);
                SCANX_TS(3);
                store_tile(s - 1);
            }
            if (ok) ok = tc::mbar_wait(epi_done, epi_rounds & 1, p.dbg, 0x1400);
            store_tile(T - 1);
            tc::tma_store_wait_all();
        }
    } else {
        // ---- epilogue.  TMEM lane quarter q = warp & 3: q < 2 rows = hi(W) of unit 32q + lane, q >= 2 rows = lo(W) of unit
        // 32(q-2) + lane; column half = warp >> 2.  After the swap with warp ^ 2 this thread owns unit j, columns [c0, c0+8).
        const int q = warp & 3, half = warp >> 2, part = q >> 1;
        const int j = (q & 1) * 32 + lane;
        const int unit = (int)c * UNITS + j;
        const int c0 = 16 * half + 8 * part;               // == 8 * (warp >> 1)
        const int tid = threadIdx.x;
        const float bhn = p.b_hn[d * H + unit];
        float hprev[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) hprev[i] = has_h0 ? p.h0[((int64_t)d * B + tile * NB + c0 + i) * H + unit] : 0.f;
        uint32_t h_off[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) h_off[i] = c * H_CHUNK + tc::sw128_offset(c0 + i, j);
        // the 16-byte piece this lane forwards to the peers: 8 units of lane group lane/8, batch row c0 + lane%8
        const uint32_t fwd_off = c * H_CHUNK + tc::sw128_offset(c0 + (lane & 7), (q & 1) * 32 + (lane >> 3) * 8);
        const uint32_t sIn_u = tc::smem_u32(sIn), sH_u = tc::smem_u32(sH);
        const uint32_t xmine = tc::smem_u32(sX) + (uint32_t)((warp * 6 * 32 + lane) * 16);
        const uint32_t xpeer = tc::smem_u32(sX) + (uint32_t)(((warp ^ 2) * 6 * 32 + lane) * 16);
        constexpr float L2E = 1.4426950408889634f;
        bool ok = true;
        uint32_t mma_rounds = 0, xch_rounds = 0;
        // publish this thread's 8 values of h (hi, lo) in operand buffer `buf`: own tile + every peer's (st.async on the
        // source-indexed barrier: the bytes travel while the other warps still compute; a bulk DSMEM copy issued by the
        // control thread after epi_done was measured slower, 0.48 vs 0.41 ms per launch), then one arrival per warp on epi_done
        auto publish = [&](const float (&h8)[8], int buf, bool to_peers) {
            const uint32_t hb = sH_u + (uint32_t)buf * 2 * TILE_BYTES;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                __nv_bfloat16 hi, lo;
                split_bf16(h8[i], hi, lo);
                tc::sts_bf16(hb + h_off[i], hi);
                tc::sts_bf16(hb + TILE_BYTES + h_off[i], lo);
            }
            tc::tcgen05_fence_before();
            if (CS > 1 && to_peers) {
                __syncwarp();
                const uint32_t a_hi = hb + fwd_off, a_lo = a_hi + TILE_BYTES, a_bar = tc::smem_u32(&h_full[buf * 4 + (int)c]);
                const uint4 vh = tc::lds_u4(a_hi);
                const uint4 vl = tc::lds_u4(a_lo);
#pragma unroll
                for (uint32_t i = 1; i < (uint32_t)CS; ++i) {
                    const uint32_t pr = (c + i) % CS;
                    const uint32_t rbar = tc::mapa_u32(a_bar, pr);
                    tc::st_async_v4(tc::mapa_u32(a_hi, pr), vh, rbar);
                    tc::st_async_v4(tc::mapa_u32(a_lo, pr), vl, rbar);
                }
            }
            tc::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(epi_done);
        };
        for (int s = 0; s < T; ++s) {
            const int t = d == 0 ? s : T - 1 - s;
            const size_t blk = blk_index(d, tile, t, (int)c, ntiles, T, CS);
            float gr[8], gz[8], gn[8];
            {
                const int st = s % NSF;
                if (ok) ok = tc::mbar_wait(&in_full[st], (s / NSF) & 1, p.dbg, 0x1200 + (s & 0xff));
                const uint32_t gp = sIn_u + (uint32_t)st * GI_BLOCK + 32u * tid;
                const float4 a0 = tc::lds_f4(gp), a1 = tc::lds_f4(gp + 16), b0 = tc::lds_f4(gp + 8192), b1 = tc::lds_f4(gp + 8192 + 16),
                             n0 = tc::lds_f4(gp + 16384), n1 = tc::lds_f4(gp + 16384 + 16);
                gr[0] = a0.x; gr[1] = a0.y; gr[2] = a0.z; gr[3] = a0.w; gr[4] = a1.x; gr[5] = a1.y; gr[6] = a1.z; gr[7] = a1.w;
                gz[0] = b0.x; gz[1] = b0.y; gz[2] = b0.z; gz[3] = b0.w; gz[4] = b1.x; gz[5] = b1.y; gz[6] = b1.z; gz[7] = b1.w;
                gn[0] = n0.x; gn[1] = n0.y; gn[2] = n0.z; gn[3] = n0.w; gn[4] = n1.x; gn[5] = n1.y; gn[6] = n1.z; gn[7] = n1.w;
            }
            float ar[8], az[8], an[8];
            if (s > 0) {
                if (tid == 0) SCANX_TS(6);
                if (ok) ok = tc::mbar_wait(mma_done, mma_rounds & 1, p.dbg, 0x1600 + (s & 0xff));
                ++mma_rounds;
                if (tid == 0) SCANX_TS(7);
                tc::tcgen05_fence_after();
                float v[3][16];
                const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(16 * half);
                tmem_ld16f(ta, v[0]); tmem_ld16f(ta + NB, v[1]); tmem_ld16f(ta + 2 * NB, v[2]);
                tmem_ld_wait_pin(v[0], v[1], v[2]);
                if (tid == 0) SCANX_TS(8);
                // hi rows keep columns [0, 8) of their half and hand [8, 16) to the lo rows' warp, and vice versa
#pragma unroll
                for (int g = 0; g < 3; ++g) {
                    tc::sts_f4(xmine + (uint32_t)((g * 2 + 0) * 512), make_float4(part ? v[g][0] : v[g][8], part ? v[g][1] : v[g][9], part ? v[g][2] : v[g][10], part ? v[g][3] : v[g][11]));
                    tc::sts_f4(xmine + (uint32_t)((g * 2 + 1) * 512), make_float4(part ? v[g][4] : v[g][12], part ? v[g][5] : v[g][13], part ? v[g][6] : v[g][14], part ? v[g][7] : v[g][15]));
                }
                // hand-over through mbarriers (release / acquire at CTA scope): this warp's half is written -> arrive on its own
                // barrier, then wait for the partner's
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&xch[warp]);
                if (ok) ok = tc::mbar_wait(&xch[warp ^ 2], xch_rounds & 1, p.dbg, 0x1900 + (s & 0xff));
                ++xch_rounds;
                float o[3][8];
#pragma unroll
                for (int g = 0; g < 3; ++g) {
                    const float4 x0 = tc::lds_f4(xpeer + (uint32_t)((g * 2 + 0) * 512)), x1 = tc::lds_f4(xpeer + (uint32_t)((g * 2 + 1) * 512));
                    o[g][0] = x0.x; o[g][1] = x0.y; o[g][2] = x0.z; o[g][3] = x0.w; o[g][4] = x1.x; o[g][5] = x1.y; o[g][6] = x1.z; o[g][7] = x1.w;
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    ar[i] = (part ? v[0][8 + i] : v[0][i]) + o[0][i];
                    az[i] = (part ? v[1][8 + i] : v[1][i]) + o[1][i];
                    an[i] = (part ? v[2][8 + i] : v[2][i]) + o[2][i];
                }
                if (tid == 0) SCANX_TS(9);
            } else if (has_h0) {
                // step 0 with an initial state: W_hh h0 comes from the caller (fp32 FFMA GEMM, [D][B][3H])
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float* gp = p.gh0 + ((int64_t)d * B + tile * NB + c0 + i) * 3 * H + unit;
                    ar[i] = gp[0]; az[i] = gp[H]; an[i] = gp[2 * H];
                }
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) { ar[i] = 0.f; az[i] = 0.f; an[i] = 0.f; }
            }
            float r8[8], z8[8], n8[8], hn8[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                // r = 1/(1+ea), z = 1/(1+eb), n = tanh(cn) = 1 - 2/(1+et): three ex2, two reciprocals
                const float ea = ex2_approx(-L2E * clampf(gr[i] + ar[i], -30.f, 30.f));
                const float r = rcp_approx(1.f + ea);
                hn8[i] = an[i] + bhn;
                const float cn = clampf(fmaf(r, hn8[i], gn[i]), -15.f, 15.f);
                const float eb = ex2_approx(-L2E * clampf(gz[i] + az[i], -30.f, 30.f));
                const float et = ex2_approx(2.f * L2E * cn);
                const float inv = rcp_approx((1.f + eb) * (1.f + et));
                const float z = inv * (1.f + et);
                const float n = fmaf(-2.f * inv, 1.f + eb, 1.f);
                r8[i] = r; z8[i] = z; n8[i] = n;
                hprev[i] = fmaf(z, hprev[i] - n, n);
            }
            if (tid == 0) SCANX_TS(10);
            publish(hprev, s & 1, s + 1 < T);
            // the ring slot is released only HERE, after this step's results (which consume every value loaded from the slot)
            // have been written: an arrive right behind the loads was seen to overtake them (the loads sat in the LSU queue behind
            // the previous step's global stores), so the producer's next bulk copy replaced the slot before it had been read
            if (lane == 0) tc::mbar_arrive(&in_empty[s % NSF]);
            if (tid == 0) SCANX_TS(11);
            {   // stash (off the chain)
                float4* gs = reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(p.GX) + blk * G_BLOCK) + 2 * tid;
                gs[0] = make_float4(r8[0], r8[1], r8[2], r8[3]); gs[1] = make_float4(r8[4], r8[5], r8[6], r8[7]);
                gs[512] = make_float4(z8[0], z8[1], z8[2], z8[3]); gs[513] = make_float4(z8[4], z8[5], z8[6], z8[7]);
                gs[1024] = make_float4(n8[0], n8[1], n8[2], n8[3]); gs[1025] = make_float4(n8[4], n8[5], n8[6], n8[7]);
                gs[1536] = make_float4(hn8[0], hn8[1], hn8[2], hn8[3]); gs[1537] = make_float4(hn8[4], hn8[5], hn8[6], hn8[7]);
                float4* ys = reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(p.YBX) + blk * YB_BLOCK) + 2 * tid;
                ys[0] = make_float4(hprev[0], hprev[1], hprev[2], hprev[3]); ys[1] = make_float4(hprev[4], hprev[5], hprev[6], hprev[7]);
            }
            if (s == T - 1 && p.hn_out) {
#pragma unroll
                for (int i = 0; i < 8; ++i) p.hn_out[((int64_t)d * B + tile * NB + c0 + i) * H + unit] = hprev[i];
            }
        }
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    if (CS > 1) tc::cluster_sync_all();
    if (warp == EPI_WARPS) tc::tmem_dealloc(tmem, 512);
}




// =================================================================================================
// Forward, ping-pong form.  The 32-row batch tile is worked as two 16-row SUB-TILES that alternate on every resource:
// while the tensor pipe multiplies sub-tile B (N = 16, 96 MMAs of 8 cycles), the epilogue warps do sub-tile A's gate math
// and the DSMEM network carries A's new h to the peers.  Measured on the single-tile kernel above: a step is 6200 cycles of
// which the MMAs are ~1540; ~1500-2600 go to the epilogue warps sitting in st.async (DSMEM takes ~12-16 B/clk per SM, the
// sender pays) and ~570 to the hi/lo lane-half swap.  Here
//   * epilogue warps with TMEM lanes 0-63 (hi rows of W, "math warps": warps 0,1,4,5) own unit j x 8 batch columns of the
//     sub-tile outright; the warps of lanes 64-127 (lo rows, warps 2,3,6,7) only hand their partial sums (W_lo h, ~2^-9 of
//     the total) over through shared memory, and then act as the SENDERS of the finished chunk (st.async, off the math warps);
//   * all external layouts stay those of the 32-row tile (the sub-tile is the thread range [128 sub, 128 sub + 128) of a
//     blocked block; per gate one 4 KB piece).
// =================================================================================================
constexpr int NBS = 16;                    // rows of a sub-tile = UMMA N
constexpr int HS_CHUNK = NBS * 128;        // [16 x 64] bf16 K-major chunk (2 KB)
constexpr int GI_SUB = 3 * 128 * 32;       // gi of one sub-tile step: 3 gates x 128 threads x 8 floats (12 KB)
constexpr int NS2 = 3;                     // gi ring depth per sub-tile
constexpr int XBUF2 = 4 * 6 * 32 * 16;     // lo -> hi partial sums of one sub-tile: [lo warp][gate*2+k][lane] float4 (12 KB)

static inline size_t fwd2_smem_bytes(int H) {
    const int KC = H / 64;
    return (size_t)2 * 2 * 2 * KC * HS_CHUNK + (size_t)2 * NS2 * GI_SUB + (size_t)2 * XBUF2 + 1024 + 1024;
}

template <int H, bool FIRST>
__device__ __forceinline__ void fwd2_issue_chunk(uint32_t tmem_d, uint32_t tmem_a_chunk, uint64_t desc_hi, uint64_t desc_lo) {
    constexpr uint32_t idesc = tc::umma_idesc_bf16(128, NBS);
#pragma unroll
    for (int g = 0; g < 3; ++g) {
#pragma unroll
        for (int part = 0; part < 2; ++part) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
                tcs::umma_bf16_ts(tmem_d + (uint32_t)(g * NBS), tmem_a_chunk + (uint32_t)(g * (H / 2) + kk * 8),
                                  (part ? desc_lo : desc_hi) + (uint64_t)(2 * kk), idesc, (FIRST && part == 0 && kk == 0) ? 0u : 1u);
        }
    }
}

__device__ __forceinline__ void pair_barrier(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void tma_store_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

template <int H>
__global__ void __launch_bounds__(THREADS, 1) gru_scanx_fwd2_kernel(const __grid_constant__ FwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int KC = H / 64, CS = KC;
    constexpr int PART_BYTES = KC * HS_CHUNK;              // one part (hi or lo) of one sub-tile operand buffer
    constexpr int SUB_BYTES = 4 * PART_BYTES;              // [2 buf][2 part]
    const int B = p.B, T = p.T;
    uint8_t* sH = smem;                                    // [2 sub][2 buf][2 part][KC][HS_CHUNK]
    uint8_t* sIn = sH + (size_t)2 * SUB_BYTES;             // [2 sub][NS2][GI_SUB]
    uint8_t* sX = sIn + (size_t)2 * NS2 * GI_SUB;          // [2 sub][XBUF2]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sX + 2 * XBUF2);
    uint64_t* h_full = bars;                 // [2 sub][2 buf][4 src]
    uint64_t* mma_done = bars + 16;          // [2 sub]
    uint64_t* epi_done = bars + 18;          // [2 sub]  one arrival per math warp
    uint64_t* in_full = bars + 20;           // [2 sub][NS2]
    uint64_t* in_empty = bars + 20 + 2 * NS2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20 + 4 * NS2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t c = CS > 1 ? tc::cluster_ctarank() : 0u;
    const int cluster_id = blockIdx.x / CS;
    const int ntiles = B / NB;
    const int d = cluster_id / ntiles, tile = cluster_id % ntiles;
    const bool has_h0 = p.h0 != nullptr && p.gh0 != nullptr;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 16; ++i) tc::mbar_init(&h_full[i], 1);
        for (int i = 0; i < 2; ++i) { tc::mbar_init(&mma_done[i], 1); tc::mbar_init(&epi_done[i], 4); }
        for (int i = 0; i < 2 * NS2; ++i) { tc::mbar_init(&in_full[i], 1); tc::mbar_init(&in_empty[i], 4); }
        // first use of every "peer chunk landed" barrier is armed before the cluster-wide sync (no early complete_tx)
        if (CS > 1)
            for (int sub = 0; sub < 2; ++sub)
                for (uint32_t u = 0; u < (uint32_t)CS; ++u) {
                    if (u == c) continue;
                    if (T > 1) tc::mbar_arrive_expect_tx(&h_full[sub * 8 + u], 2 * HS_CHUNK);         // h_0
                    if (T > 2) tc::mbar_arrive_expect_tx(&h_full[sub * 8 + 4 + u], 2 * HS_CHUNK);     // h_1
                }
        tc::fence_mbar_init();
    }
    if (warp == EPI_WARPS) tc::tmem_alloc(tmem_slot, 512);
    tc::tcgen05_fence_before();
    __syncthreads();
    if (CS > 1) tc::cluster_sync_all();
    tc::tcgen05_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (warp < EPI_WARPS) tcs::load_weights_to_tmem(p.Wimg + ((size_t)d * CS + c) * 128 * 3 * H, 3 * H, tmem, A_COL, warp, lane);
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();

    if (warp == EPI_WARPS + 1) {
        // ---- input prefetch: per sub-tile step three 4 KB pieces (one per gate) of the 32-row block
        if (tc::elect_one()) {
            bool ok = true;
            for (int s = 0; s < T; ++s) {
                const int st = s % NS2;
                const int t = d == 0 ? s : T - 1 - s;
                const uint8_t* blk = reinterpret_cast<const uint8_t*>(p.giX) + blk_index(d, tile, t, (int)c, ntiles, T, CS) * GI_BLOCK;
                for (int sub = 0; sub < 2; ++sub) {
                    uint64_t* full = &in_full[sub * NS2 + st];
                    if (s >= NS2 && ok) ok = tc::mbar_wait(&in_empty[sub * NS2 + st], ((s / NS2) - 1) & 1, p.dbg, 0x1300 + (s & 0xff));
                    uint8_t* dst = sIn + (size_t)(sub * NS2 + st) * GI_SUB;
                    tc::mbar_arrive_expect_tx(full, GI_SUB);
#pragma unroll
                    for (int g = 0; g < 3; ++g) tc::bulk_g2s(dst + g * 4096, blk + g * 8192 + sub * 4096, 4096, full);
                }
            }
        }
    } else if (warp == EPI_WARPS) {
        // ---- control thread: alternates the two sub-tiles
        if (tc::elect_one()) {
            bool ok = true;
            uint32_t epi_rounds[2] = {0u, 0u}, hf_use[4] = {0u, 0u, 0u, 0u};
            auto store_tile = [&](int sub, int step) {
                const int tt = d == 0 ? step : T - 1 - step;
                const uint8_t* src = sH + (size_t)sub * SUB_BYTES + (size_t)(step & 1) * 2 * PART_BYTES + (size_t)c * HS_CHUNK;
                tc::tma_store_2d(&p.tmYhi, src, d * H + (int)c * UNITS, tt * B + tile * NB + sub * NBS);
                tc::tma_store_2d(&p.tmYlo, src + PART_BYTES, d * H + (int)c * UNITS, tt * B + tile * NB + sub * NBS);
                tc::tma_store_commit();
            };
            const uint32_t hb0 = tc::smem_u32(sH);
            for (int s = 1; s < T; ++s) {
                const int pb = (s - 1) & 1;
#pragma unroll
                for (int sub = 0; sub < 2; ++sub) {
                    const uint32_t tb = hb0 + (uint32_t)sub * SUB_BYTES + (uint32_t)pb * 2 * PART_BYTES;
                    const uint32_t td = tmem + (uint32_t)(sub * 48);
                    if (ok) ok = tc::mbar_wait(&epi_done[sub], epi_rounds[sub] & 1, p.dbg, 0x1400 + (s & 0xff));
                    ++epi_rounds[sub];
                    if (sub == 0) SCANX_TS(0);
                    tc::tcgen05_fence_after();
                    fwd2_issue_chunk<H, true>(td, tmem + A_COL + c * 32, tc::umma_desc_k_sw128(tb + c * HS_CHUNK),
                                              tc::umma_desc_k_sw128(tb + PART_BYTES + c * HS_CHUNK));
                    if (sub == 0) SCANX_TS(1);
                    for (uint32_t i = 1; i < (uint32_t)CS; ++i) {
                        const uint32_t u = (c + i) % CS;
                        uint64_t* hf = &h_full[sub * 8 + pb * 4 + u];
                        if (ok) ok = tc::mbar_wait(hf, hf_use[sub * 2 + pb] & 1, p.dbg, 0x1500 + (s & 0xff));
                        if (s + 2 < T) tc::mbar_arrive_expect_tx(hf, 2 * HS_CHUNK);
                        tc::tcgen05_fence_after();
                        fwd2_issue_chunk<H, false>(td, tmem + A_COL + u * 32, tc::umma_desc_k_sw128(tb + u * HS_CHUNK),
                                                   tc::umma_desc_k_sw128(tb + PART_BYTES + u * HS_CHUNK));
                    }
                    ++hf_use[sub * 2 + pb];
                    tma_store_wait_read1();               // the tile of this sub-tile stored two steps ago has been read
                    tc::umma_commit(&mma_done[sub]);
                    if (sub == 0) SCANX_TS(3);
                    store_tile(sub, s - 1);
                }
            }
            for (int sub = 0; sub < 2; ++sub) {
                if (ok) ok = tc::mbar_wait(&epi_done[sub], epi_rounds[sub] & 1, p.dbg, 0x1400);
                store_tile(sub, T - 1);
            }
            tc::tma_store_wait_all();
        }
    } else {
        const int q = warp & 3, half = warp >> 2;
        const bool math = q < 2;                                   // TMEM lanes 0-63: hi rows
        const int j = (q & 1) * 32 + lane;
        const int unit = (int)c * UNITS + j;
        const int pair_id = 2 + (q & 1) + 2 * half;
        const uint32_t sIn_u = tc::smem_u32(sIn), sH_u = tc::smem_u32(sH), sX_u = tc::smem_u32(sX);
        const int lw = (q & 1) + 2 * half;                         // index of this warp among the 4 lo (or 4 math) warps
        const uint32_t xoff = (uint32_t)((lw * 6 * 32 + lane) * 16);
        uint32_t mma_rounds[2] = {0u, 0u};
        bool ok = true;
        if (math) {
            // ---- math warps: unit j, batch columns [8*half, +8) of each 16-row sub-tile
            const int tl = j + 64 * half;                          // thread index inside the sub-tile's half of a blocked block
            const float bhn = p.b_hn[d * H + unit];
            float hprev[2][8];
#pragma unroll
            for (int sub = 0; sub < 2; ++sub)
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    hprev[sub][i] = has_h0 ? p.h0[((int64_t)d * B + tile * NB + sub * NBS + 8 * half + i) * H + unit] : 0.f;
            uint32_t h_off[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) h_off[i] = c * HS_CHUNK + tc::sw128_offset(8 * half + i, j);
            constexpr float L2E = 1.4426950408889634f;
            for (int s = 0; s < T; ++s) {
                const int t = d == 0 ? s : T - 1 - s;
                const size_t blk = blk_index(d, tile, t, (int)c, ntiles, T, CS);
#pragma unroll
                for (int sub = 0; sub < 2; ++sub) {
                    float gr[8], gz[8], gn[8];
                    {
                        const int st = s % NS2;
                        if (ok) ok = tc::mbar_wait(&in_full[sub * NS2 + st], (s / NS2) & 1, p.dbg, 0x1200 + (s & 0xff));
                        const uint32_t gp = sIn_u + (uint32_t)(sub * NS2 + st) * GI_SUB + 32u * tl;
                        const float4 a0 = tc::lds_f4(gp), a1 = tc::lds_f4(gp + 16), b0 = tc::lds_f4(gp + 4096), b1 = tc::lds_f4(gp + 4096 + 16),
                                     n0 = tc::lds_f4(gp + 8192), n1 = tc::lds_f4(gp + 8192 + 16);
                        gr[0] = a0.x; gr[1] = a0.y; gr[2] = a0.z; gr[3] = a0.w; gr[4] = a1.x; gr[5] = a1.y; gr[6] = a1.z; gr[7] = a1.w;
                        gz[0] = b0.x; gz[1] = b0.y; gz[2] = b0.z; gz[3] = b0.w; gz[4] = b1.x; gz[5] = b1.y; gz[6] = b1.z; gz[7] = b1.w;
                        gn[0] = n0.x; gn[1] = n0.y; gn[2] = n0.z; gn[3] = n0.w; gn[4] = n1.x; gn[5] = n1.y; gn[6] = n1.z; gn[7] = n1.w;
                    }
                    float ar[8], az[8], an[8];
                    if (s > 0) {
                        if (threadIdx.x == 0 && sub == 0) SCANX_TS(6);
                        if (ok) ok = tc::mbar_wait(&mma_done[sub], mma_rounds[sub] & 1, p.dbg, 0x1600 + (s & 0xff));
                        ++mma_rounds[sub];
                        if (threadIdx.x == 0 && sub == 0) SCANX_TS(7);
                        tc::tcgen05_fence_after();
                        const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(sub * 48 + 8 * half);
                        tcs::tmem_ld8(ta, ar); tcs::tmem_ld8(ta + NBS, az); tcs::tmem_ld8(ta + 2 * NBS, an);
                        tc::tmem_ld_wait();
                        if (threadIdx.x == 0 && sub == 0) SCANX_TS(8);
                        pair_barrier(pair_id);                     // the lo rows' partial sums of this sub-tile are in shared memory
                        const uint32_t xp = sX_u + (uint32_t)sub * XBUF2 + xoff;
#pragma unroll
                        for (int g = 0; g < 3; ++g) {
                            const float4 x0 = tc::lds_f4(xp + (uint32_t)((g * 2 + 0) * 512)), x1 = tc::lds_f4(xp + (uint32_t)((g * 2 + 1) * 512));
                            float* a = g == 0 ? ar : (g == 1 ? az : an);
                            a[0] += x0.x; a[1] += x0.y; a[2] += x0.z; a[3] += x0.w; a[4] += x1.x; a[5] += x1.y; a[6] += x1.z; a[7] += x1.w;
                        }
                        if (threadIdx.x == 0 && sub == 0) SCANX_TS(9);
                    } else if (has_h0) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float* gp0 = p.gh0 + ((int64_t)d * B + tile * NB + sub * NBS + 8 * half + i) * 3 * H + unit;
                            ar[i] = gp0[0]; az[i] = gp0[H]; an[i] = gp0[2 * H];
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) { ar[i] = 0.f; az[i] = 0.f; an[i] = 0.f; }
                    }
                    float r8[8], z8[8], n8[8], hn8[8];
                    const uint32_t hb = sH_u + (uint32_t)sub * SUB_BYTES + (uint32_t)(s & 1) * 2 * PART_BYTES;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float ea = ex2_approx(-L2E * clampf(gr[i] + ar[i], -30.f, 30.f));
                        const float r = rcp_approx(1.f + ea);
                        hn8[i] = an[i] + bhn;
                        const float cn = clampf(fmaf(r, hn8[i], gn[i]), -15.f, 15.f);
                        const float eb = ex2_approx(-L2E * clampf(gz[i] + az[i], -30.f, 30.f));
                        const float et = ex2_approx(2.f * L2E * cn);
                        const float inv = rcp_approx((1.f + eb) * (1.f + et));
                        const float z = inv * (1.f + et);
                        const float n = fmaf(-2.f * inv, 1.f + eb, 1.f);
                        r8[i] = r; z8[i] = z; n8[i] = n;
                        const float h = fmaf(z, hprev[sub][i] - n, n);
                        hprev[sub][i] = h;
                        __nv_bfloat16 hi, lo;
                        split_bf16(h, hi, lo);
                        tc::sts_bf16(hb + h_off[i], hi);
                        tc::sts_bf16(hb + PART_BYTES + h_off[i], lo);
                    }
                    if (threadIdx.x == 0 && sub == 0) SCANX_TS(10);
                    tc::tcgen05_fence_before();
                    tc::fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(&epi_done[sub]);
                    // ring slot released only after the publish (which consumes every value loaded from it), see tc_scan.cuh
                    if (lane == 0) tc::mbar_arrive(&in_empty[sub * NS2 + s % NS2]);
                    if (threadIdx.x == 0 && sub == 0) SCANX_TS(11);
                    {   // stash (off the chain): thread tl + 128 sub of the 32-row block
                        float4* gs = reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(p.GX) + blk * G_BLOCK) + 2 * (tl + 128 * sub);
                        gs[0] = make_float4(r8[0], r8[1], r8[2], r8[3]); gs[1] = make_float4(r8[4], r8[5], r8[6], r8[7]);
                        gs[512] = make_float4(z8[0], z8[1], z8[2], z8[3]); gs[513] = make_float4(z8[4], z8[5], z8[6], z8[7]);
                        gs[1024] = make_float4(n8[0], n8[1], n8[2], n8[3]); gs[1025] = make_float4(n8[4], n8[5], n8[6], n8[7]);
                        gs[1536] = make_float4(hn8[0], hn8[1], hn8[2], hn8[3]); gs[1537] = make_float4(hn8[4], hn8[5], hn8[6], hn8[7]);
                        float4* ys = reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(p.YBX) + blk * YB_BLOCK) + 2 * (tl + 128 * sub);
                        ys[0] = make_float4(hprev[sub][0], hprev[sub][1], hprev[sub][2], hprev[sub][3]);
                        ys[1] = make_float4(hprev[sub][4], hprev[sub][5], hprev[sub][6], hprev[sub][7]);
                    }
                    if (s == T - 1 && p.hn_out) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) p.hn_out[((int64_t)d * B + tile * NB + sub * NBS + 8 * half + i) * H + unit] = hprev[sub][i];
                    }
                }
            }
        } else {
            // ---- lo-row warps: hand W_lo h over to the math warps, then forward the finished chunk of h to the peers
            uint32_t epi_rounds[2] = {0u, 0u};
            // this lane's 16-byte piece of the [16 x 64] chunk (hi and lo): row = piece / 8, units 8*(piece % 8)..+7
            const int piece = lw * 32 + lane;
            const uint32_t piece_off = c * HS_CHUNK + tc::sw128_offset(piece >> 3, (piece & 7) * 8);
            for (int s = 0; s < T; ++s) {
                // both hand-overs first (the math warps wait for them), then both forwards (this warp sits in st.async for
                // ~1000 cycles per sub-tile: DSMEM takes 12-16 B/clk and the sender pays)
                if (s > 0) {
#pragma unroll
                    for (int sub = 0; sub < 2; ++sub) {
                        if (ok) ok = tc::mbar_wait(&mma_done[sub], mma_rounds[sub] & 1, p.dbg, 0x1700 + (s & 0xff));
                        ++mma_rounds[sub];
                        tc::tcgen05_fence_after();
                        float v[3][8];
                        const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(sub * 48 + 8 * half);
                        tcs::tmem_ld8(ta, v[0]); tcs::tmem_ld8(ta + NBS, v[1]); tcs::tmem_ld8(ta + 2 * NBS, v[2]);
                        tc::tmem_ld_wait();
                        tc::tcgen05_fence_before();
                        const uint32_t xp = sX_u + (uint32_t)sub * XBUF2 + xoff;
#pragma unroll
                        for (int g = 0; g < 3; ++g) {
                            tc::sts_f4(xp + (uint32_t)((g * 2 + 0) * 512), make_float4(v[g][0], v[g][1], v[g][2], v[g][3]));
                            tc::sts_f4(xp + (uint32_t)((g * 2 + 1) * 512), make_float4(v[g][4], v[g][5], v[g][6], v[g][7]));
                        }
                        pair_barrier(pair_id);
                    }
                }
#pragma unroll
                for (int sub = 0; sub < 2; ++sub) {
                    if (CS > 1 && s + 1 < T) {
                        if (ok) ok = tc::mbar_wait(&epi_done[sub], epi_rounds[sub] & 1, p.dbg, 0x1800 + (s & 0xff));
                        const int buf = s & 1;
                        const uint32_t a_hi = sH_u + (uint32_t)sub * SUB_BYTES + (uint32_t)buf * 2 * PART_BYTES + piece_off, a_lo = a_hi + PART_BYTES;
                        const uint32_t a_bar = tc::smem_u32(&h_full[sub * 8 + buf * 4 + (int)c]);
                        const uint4 vh = tc::lds_u4(a_hi);
                        const uint4 vl = tc::lds_u4(a_lo);
#pragma unroll
                        for (uint32_t i = 1; i < (uint32_t)CS; ++i) {
                            const uint32_t pr = (c + i) % CS;
                            const uint32_t rbar = tc::mapa_u32(a_bar, pr);
                            tc::st_async_v4(tc::mapa_u32(a_hi, pr), vh, rbar);
                            tc::st_async_v4(tc::mapa_u32(a_lo, pr), vl, rbar);
                        }
                    }
                    ++epi_rounds[sub];
                }
            }
        }
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    if (CS > 1) tc::cluster_sync_all();
    if (warp == EPI_WARPS) tc::tmem_dealloc(tmem, 512);
}



static inline cudaError_t launch_fwd(const FwdParams& p_in, cudaStream_t st) {
    FwdParams p = p_in;
    if ((p.H != 128 && p.H != 256) || p.B % NB != 0) return cudaErrorInvalidValue;
    // H = 256: the ping-pong form (two 16-row sub-tiles alternate on tensor pipe / epilogue / DSMEM), 10 % faster per step;
    // BIGRU_X3_FWD=single selects the single-tile kernel
    static const bool single = [] { const char* e = getenv("BIGRU_X3_FWD"); return e && e[0] == 's'; }();
    const bool pp = p.H == 256 && !single;
    {
        const uint64_t dims[2] = {(uint64_t)p.D * p.H, (uint64_t)p.T * p.B};
        const uint64_t strides[1] = {(uint64_t)p.D * p.H * 2};
        const uint32_t box[2] = {64u, (uint32_t)(pp ? NBS : NB)};          // the ping-pong form stores 16-row sub-tiles
        if (make_tmap_bf16(&p.tmYhi, p.Yhi, 2, dims, strides, box) != 0 || make_tmap_bf16(&p.tmYlo, p.Ylo, 2, dims, strides, box) != 0)
            return cudaErrorInvalidValue;
    }
    const int CS = p.H / UNITS;
    const size_t smem = pp ? fwd2_smem_bytes(p.H) : fwd_smem_bytes(p.H);
    void (*kern)(FwdParams) = p.H == 128 ? gru_scanx_fwd_kernel<128> : (pp ? gru_scanx_fwd2_kernel<256> : gru_scanx_fwd_kernel<256>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(p.D * (p.B / NB) * CS));
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, p);
}

// =================================================================================================
// Backward scan (BPTT), reduction-partitioned (see the header comment).
//   A operand (TMEM): lane i of row block rb = part*NKH + kh holds split_part(W_hh[q][k]) for k = 128*kh + i and the
//   CTA's own gate rows q = g*H + 64c + jj, K index kq = g*64 + jj (192 = 12 K-steps): image [D][CS][128 lanes][NRB*192].
//   B operand (smem): this CTA's dgh tile [32 x 192] (da_r | da_z | da_n*r of its 64 units), hi and lo.
//   D[rb] (32 columns each): partial dh for output unit k; hi + lo blocks of one k are summed in registers, partials of
//   units owned by another CTA travel to its receive buffer, the owner adds the CS contributions.
// =================================================================================================
struct BwdParams {
    int B, T, H, D;
    const __nv_bfloat16* WTimg;
    const float* GX;
    const float* YBX;
    const float* dYBX;              // lower layers
    const float* h0;                // nullable [D][B][H]: h_prev of the first forward step
    const float* dlogits;           // top layer (see tc_scan.cuh)
    const float* lin_w;
    const int* arg;
    int C;
    __nv_bfloat16 *dgi_hi, *dgi_lo;     // [R][D*3H]
    __nv_bfloat16 *dghn_hi, *dghn_lo;   // [R][D*H]
    CUtensorMap tmGIh, tmGIl, tmGNh, tmGNl;
    float* db_ih;
    float* db_hh;
    int64_t dir_stride;
    float* dh0;                     // nullable [D][B][H]: gradient of the initial hidden state
    unsigned int* dbg;
#ifdef BIGRU_SCAN_TIMING
    unsigned long long* ts;
#endif
};

// both parts (hi, lo) of the dgh tile against row block `rb`: 24 MMAs
__device__ __forceinline__ void bwd_issue_block(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_hi, uint64_t desc_lo) {
    constexpr uint32_t idesc = tc::umma_idesc_bf16(128, NB);
#pragma unroll
    for (int part = 0; part < 2; ++part) {
#pragma unroll
        for (int g = 0; g < 3; ++g) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
                tcs::umma_bf16_ts(tmem_d, tmem_a + (uint32_t)((g * 4 + kk) * 8),
                                  (part ? desc_lo : desc_hi) + (uint64_t)(g * (H_CHUNK >> 4) + 2 * kk), idesc,
                                  (part == 0 && g == 0 && kk == 0) ? 0u : 1u);
        }
    }
}

template <int H>
__global__ void __launch_bounds__(THREADS, 1) gru_scanx_bwd_kernel(const __grid_constant__ BwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int CS = H / 64, NKH = H / 128, NRB = 2 * NKH;
    constexpr int DT_BYTES = 3 * H_CHUNK;                 // one part of one dgh tile
    const int B = p.B, T = p.T;
    uint8_t* sD = smem;                                    // [2 buf][2 part][3 gates][H_CHUNK]
    uint8_t* sN = sD + (size_t)4 * DT_BYTES;               // [2 buf][2 part][H_CHUNK]   da_n (dgi n-gate rows, store only)
    uint8_t* sR = sN + (size_t)4 * H_CHUNK;                // [2 buf][4 src][8 cg][64 j] float4
    uint8_t* sIn = sR + (size_t)2 * RECV_BYTES;            // [NSB][G | YB | dY]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sIn + (size_t)NSB * BWD_STAGE);
    uint64_t* recv_full = bars;        // [2]
    uint64_t* mma_a = bars + 2;        // row blocks of the other k half done
    uint64_t* mma_b = bars + 3;        // all row blocks done
    uint64_t* epi_done = bars + 4;
    uint64_t* st_done = bars + 5;
    uint64_t* in_full = bars + 6;      // [NSB]
    uint64_t* in_empty = bars + 6 + NSB;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6 + 2 * NSB);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t c = tc::cluster_ctarank();
    const int cluster_id = blockIdx.x / CS;
    const int ntiles = B / NB;
    const int d = cluster_id / ntiles, tile = cluster_id % ntiles;
    const bool top = p.dlogits != nullptr;
    const int kh_own = (int)c >> 1;                        // k half that contains this CTA's own units (H = 256)

    if (threadIdx.x == 0) {
        tc::mbar_init(&recv_full[0], 1);
        tc::mbar_init(&recv_full[1], 1);
        tc::mbar_init(mma_a, 1);
        tc::mbar_init(mma_b, 1);
        tc::mbar_init(epi_done, EPI_WARPS);
        tc::mbar_init(st_done, EPI_WARPS);
        for (int i = 0; i < NSB; ++i) { tc::mbar_init(&in_full[i], 1); tc::mbar_init(&in_empty[i], EPI_WARPS); }
        tc::fence_mbar_init();
    }
    if (warp == EPI_WARPS) tc::tmem_alloc(tmem_slot, 512);
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::cluster_sync_all();
    tc::tcgen05_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (warp < EPI_WARPS)
        tcs::load_weights_to_tmem(p.WTimg + ((size_t)d * CS + c) * 128 * (NRB * 192), NRB * 192, tmem, A_COL, warp, lane);
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();

    if (warp == EPI_WARPS + 1) {
        if (tc::elect_one()) {
            bool ok = true;
            for (int s = 0; s < T; ++s) {
                const int st = s % NSB;
                if (s >= NSB && ok) ok = tc::mbar_wait(&in_empty[st], ((s / NSB) - 1) & 1, p.dbg, 0x2300 + (s & 0xff));
                const int t = d == 0 ? T - 1 - s : s;
                const bool first = d == 0 ? t == 0 : t == T - 1;
                uint8_t* dst = sIn + (size_t)st * BWD_STAGE;
                tc::mbar_arrive_expect_tx(&in_full[st], (uint32_t)(G_BLOCK + (top ? 0 : DY_BLOCK) + (first ? 0 : YB_BLOCK)));
                const size_t blk = blk_index(d, tile, t, (int)c, ntiles, T, CS);
                tc::bulk_g2s(dst, reinterpret_cast<const uint8_t*>(p.GX) + blk * G_BLOCK, G_BLOCK, &in_full[st]);
                if (!top) tc::bulk_g2s(dst + G_BLOCK + YB_BLOCK, reinterpret_cast<const uint8_t*>(p.dYBX) + blk * DY_BLOCK, DY_BLOCK, &in_full[st]);
                if (!first) {
                    const size_t pblk = blk_index(d, tile, d == 0 ? t - 1 : t + 1, (int)c, ntiles, T, CS);
                    tc::bulk_g2s(dst + G_BLOCK, reinterpret_cast<const uint8_t*>(p.YBX) + pblk * YB_BLOCK, YB_BLOCK, &in_full[st]);
                }
            }
        }
    } else if (warp == EPI_WARPS) {
        if (tc::elect_one()) {
            bool ok = true;
            auto store_tile = [&](int step) {
                const int tt = d == 0 ? T - 1 - step : step;
                const int row = tt * B + tile * NB;
                const uint8_t* tb = sD + (size_t)(step & 1) * 2 * DT_BYTES;
                const uint8_t* nb = sN + (size_t)(step & 1) * 2 * H_CHUNK;
                const int cu = (int)c * UNITS;
                tc::tma_store_2d(&p.tmGIh, tb, d * 3 * H + cu, row);                               // da_r
                tc::tma_store_2d(&p.tmGIl, tb + DT_BYTES, d * 3 * H + cu, row);
                tc::tma_store_2d(&p.tmGIh, tb + H_CHUNK, d * 3 * H + H + cu, row);                 // da_z
                tc::tma_store_2d(&p.tmGIl, tb + DT_BYTES + H_CHUNK, d * 3 * H + H + cu, row);
                tc::tma_store_2d(&p.tmGIh, nb, d * 3 * H + 2 * H + cu, row);                       // da_n
                tc::tma_store_2d(&p.tmGIl, nb + H_CHUNK, d * 3 * H + 2 * H + cu, row);
                tc::tma_store_2d(&p.tmGNh, tb + 2 * H_CHUNK, d * H + cu, row);                     // da_n * r
                tc::tma_store_2d(&p.tmGNl, tb + DT_BYTES + 2 * H_CHUNK, d * H + cu, row);
                tc::tma_store_commit();
            };
            const uint32_t db0 = tc::smem_u32(sD);
            const int Tend = T + (p.dh0 ? 1 : 0);          // one more product (no gate math) when d(h0) is wanted
            for (int s = 1; s < Tend; ++s) {
                const int pb = (s - 1) & 1;
                if (ok) ok = tc::mbar_wait(epi_done, (s - 1) & 1, p.dbg, 0x2700 + (s & 0xff));
                SCANX_TS(0);
                tc::tcgen05_fence_after();
                tc::mbar_arrive_expect_tx(&recv_full[s & 1], (uint32_t)(CS - 1) * 8192u);
                const uint64_t dhi = tc::umma_desc_k_sw128(db0 + (uint32_t)pb * 2 * DT_BYTES);
                const uint64_t dlo = tc::umma_desc_k_sw128(db0 + (uint32_t)pb * 2 * DT_BYTES + DT_BYTES);
                if (NKH == 2) {
                    const int ko = 1 - kh_own;
                    bwd_issue_block(tmem + (uint32_t)((0 * NKH + ko) * NB), tmem + A_COL + (uint32_t)((0 * NKH + ko) * 96), dhi, dlo);
                    bwd_issue_block(tmem + (uint32_t)((1 * NKH + ko) * NB), tmem + A_COL + (uint32_t)((1 * NKH + ko) * 96), dhi, dlo);
                    tc::umma_commit(mma_a);
                }
                SCANX_TS(1);
                {
                    const int ko = NKH == 2 ? kh_own : 0;
                    bwd_issue_block(tmem + (uint32_t)((0 * NKH + ko) * NB), tmem + A_COL + (uint32_t)((0 * NKH + ko) * 96), dhi, dlo);
                    bwd_issue_block(tmem + (uint32_t)((1 * NKH + ko) * NB), tmem + A_COL + (uint32_t)((1 * NKH + ko) * 96), dhi, dlo);
                }
                tc::tma_store_wait_read();
                tc::umma_commit(mma_b);
                SCANX_TS(3);
                if (ok) ok = tc::mbar_wait(st_done, (s - 1) & 1, p.dbg, 0x2a00 + (s & 0xff));
                store_tile(s - 1);
            }
            if (Tend == T) {
                if (ok) ok = tc::mbar_wait(epi_done, (T - 1) & 1, p.dbg, 0x2700);
                if (ok) ok = tc::mbar_wait(st_done, (T - 1) & 1, p.dbg, 0x2a00);
                store_tile(T - 1);
            }
            tc::tma_store_wait_all();
        }
    } else {
        // ---- epilogue.  Owner role: unit j = (warp & 1)*32 + lane of this CTA, batch columns [8*(warp >> 1), +8).
        //      Partial-sum role: TMEM lane quarter q = warp & 3 -> output unit k = 128*kh + 32q + lane, columns [16*half, +16).
        const int q = warp & 3, half = warp >> 2;
        const int j = (warp & 1) * 32 + lane;
        const int unit = (int)c * UNITS + j;
        const int c0 = 8 * (warp >> 1);
        const int tid = threadIdx.x;
        float dhz[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) dhz[i] = 0.f;
        float h_avg[8], h_max[8];
        int h_arg[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { h_avg[i] = 0.f; h_max[i] = 0.f; h_arg[i] = -1; }
        if (top) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int b = tile * NB + c0 + i;
                float dl = 0.f, dm = 0.f, da = 0.f;
                for (int cc = 0; cc < p.C; ++cc) {
                    const float g = p.dlogits[(int64_t)b * p.C + cc];
                    const float* w = p.lin_w + (int64_t)cc * 3 * H;
                    dl = fmaf(g, w[unit], dl); dm = fmaf(g, w[H + unit], dm); da = fmaf(g, w[2 * H + unit], da);
                }
                dhz[i] = dl;
                h_avg[i] = da / (float)T; h_max[i] = dm; h_arg[i] = p.arg[(int64_t)b * H + unit];
            }
        }
        float sb_r = 0.f, sb_z = 0.f, sb_n = 0.f, sb_nr = 0.f;
        uint32_t e_off[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) e_off[i] = tc::sw128_offset(c0 + i, j);
        // partial-sum destination inside a receive buffer: [src = c][cg = 4*half + i][jd] float4, jd = (q & 1)*32 + lane
        const uint32_t r_off = (((uint32_t)c * 8 + 4 * half) * 64 + (uint32_t)((q & 1) * 32 + lane)) * 16;
        const uint32_t sIn_u = tc::smem_u32(sIn), sR_u = tc::smem_u32(sR), sD_u = tc::smem_u32(sD), sN_u = tc::smem_u32(sN);
        bool ok = true;
        // recurrent part of dh for this thread's (unit, 8 columns) at step s: every CTA's row blocks -> partial sums (hi + lo
        // blocks of one k share a lane) -> owner's receive buffer (st.async / local store) -> sum over the CS sources
        auto reduce_partials = [&](int s, float (&acc)[8]) {
            const int buf = s & 1;
            const uint32_t rb_local = sR_u + (uint32_t)buf * RECV_BYTES;
            const uint32_t rbar_l = tc::smem_u32(&recv_full[buf]);
            auto route = [&](int kh) {
                float vh[16], vl[16];
                const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(16 * half);
                tmem_ld16f(ta + (uint32_t)((0 * NKH + kh) * NB), vh);
                tmem_ld16f(ta + (uint32_t)((1 * NKH + kh) * NB), vl);
                tmem_ld_wait_pin(vh, vl);
                const uint32_t dest = (uint32_t)(2 * kh + (q >> 1));
                const uint32_t lp = rb_local + r_off;
                if (dest == c) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        tc::sts_f4(lp + (uint32_t)(i * 64 * 16),
                                   make_float4(vh[4 * i] + vl[4 * i], vh[4 * i + 1] + vl[4 * i + 1], vh[4 * i + 2] + vl[4 * i + 2], vh[4 * i + 3] + vl[4 * i + 3]));
                } else {
                    const uint32_t ra = tc::mapa_u32(lp, dest), rb = tc::mapa_u32(rbar_l, dest);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        uint4 u;
                        u.x = __float_as_uint(vh[4 * i] + vl[4 * i]); u.y = __float_as_uint(vh[4 * i + 1] + vl[4 * i + 1]);
                        u.z = __float_as_uint(vh[4 * i + 2] + vl[4 * i + 2]); u.w = __float_as_uint(vh[4 * i + 3] + vl[4 * i + 3]);
                        tc::st_async_v4(ra + (uint32_t)(i * 64 * 16), u, rb);
                    }
                }
            };
            if (NKH == 2) {
                if (tid == 0) SCANX_TS(4);
                if (ok) ok = tc::mbar_wait(mma_a, (s - 1) & 1, p.dbg, 0x2800 + (s & 0xff));
                tc::tcgen05_fence_after();
                route(1 - kh_own);
            }
            if (ok) ok = tc::mbar_wait(mma_b, (s - 1) & 1, p.dbg, 0x2900 + (s & 0xff));
            if (tid == 0) SCANX_TS(7);
            tc::tcgen05_fence_after();
            route(NKH == 2 ? kh_own : 0);
            tc::tcgen05_fence_before();
            epi_barrier();                                   // this CTA's own contributions are in the buffer
            if (ok) ok = tc::mbar_wait_cluster(&recv_full[buf], ((s - 1) >> 1) & 1, p.dbg, 0x2b00 + (s & 0xff));
            if (tid == 0) SCANX_TS(8);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
            for (int src = 0; src < CS; ++src) {
                const uint32_t rp = rb_local + (uint32_t)((((src * 8 + 2 * (warp >> 1)) * 64) + j) * 16);
                const float4 x0 = tc::lds_f4(rp), x1 = tc::lds_f4(rp + 64 * 16);
                acc[0] += x0.x; acc[1] += x0.y; acc[2] += x0.z; acc[3] += x0.w; acc[4] += x1.x; acc[5] += x1.y; acc[6] += x1.z; acc[7] += x1.w;
            }
        };
        for (int s = 0; s < T; ++s) {
            const int t = d == 0 ? T - 1 - s : s;
            const bool first = d == 0 ? t == 0 : t == T - 1;
            float vr[8], vz[8], vn[8], vhn[8], vhp[8], vdy[8];
            {
                const int st = s % NSB;
                if (ok) ok = tc::mbar_wait(&in_full[st], (s / NSB) & 1, p.dbg, 0x2200 + (s & 0xff));
                const uint32_t gp = sIn_u + (uint32_t)st * BWD_STAGE + 32u * tid;
                const float4 a0 = tc::lds_f4(gp), a1 = tc::lds_f4(gp + 16), b0 = tc::lds_f4(gp + 8192), b1 = tc::lds_f4(gp + 8192 + 16),
                             n0 = tc::lds_f4(gp + 16384), n1 = tc::lds_f4(gp + 16384 + 16), m0 = tc::lds_f4(gp + 24576), m1 = tc::lds_f4(gp + 24576 + 16);
                float4 p0 = make_float4(0.f, 0.f, 0.f, 0.f), p1 = p0, y0 = p0, y1 = p0;
                if (!first) { p0 = tc::lds_f4(gp + G_BLOCK); p1 = tc::lds_f4(gp + G_BLOCK + 16); }
                if (!top) { y0 = tc::lds_f4(gp + G_BLOCK + YB_BLOCK); y1 = tc::lds_f4(gp + G_BLOCK + YB_BLOCK + 16); }
                vr[0] = a0.x; vr[1] = a0.y; vr[2] = a0.z; vr[3] = a0.w; vr[4] = a1.x; vr[5] = a1.y; vr[6] = a1.z; vr[7] = a1.w;
                vz[0] = b0.x; vz[1] = b0.y; vz[2] = b0.z; vz[3] = b0.w; vz[4] = b1.x; vz[5] = b1.y; vz[6] = b1.z; vz[7] = b1.w;
                vn[0] = n0.x; vn[1] = n0.y; vn[2] = n0.z; vn[3] = n0.w; vn[4] = n1.x; vn[5] = n1.y; vn[6] = n1.z; vn[7] = n1.w;
                vhn[0] = m0.x; vhn[1] = m0.y; vhn[2] = m0.z; vhn[3] = m0.w; vhn[4] = m1.x; vhn[5] = m1.y; vhn[6] = m1.z; vhn[7] = m1.w;
                vhp[0] = p0.x; vhp[1] = p0.y; vhp[2] = p0.z; vhp[3] = p0.w; vhp[4] = p1.x; vhp[5] = p1.y; vhp[6] = p1.z; vhp[7] = p1.w;
                vdy[0] = y0.x; vdy[1] = y0.y; vdy[2] = y0.z; vdy[3] = y0.w; vdy[4] = y1.x; vdy[5] = y1.y; vdy[6] = y1.z; vdy[7] = y1.w;
                if (first && p.h0) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) vhp[i] = p.h0[((int64_t)d * B + tile * NB + c0 + i) * H + unit];
                }
                if (top) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) vdy[i] = h_avg[i] + (h_arg[i] == t ? h_max[i] : 0.f);
                }
            }
            float c_n[8], c_r[8], c_z[8], pre[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float r = vr[i], z = vz[i], n = vn[i];
                c_n[i] = (1.f - z) * (1.f - n * n);
                c_r[i] = vhn[i] * r * (1.f - r);
                c_z[i] = (vhp[i] - n) * z * (1.f - z);
                pre[i] = dhz[i] + vdy[i];
            }
            float acc[8];
            const int buf = s & 1;
            if (s > 0) reduce_partials(s, acc);
            else {
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] = 0.f;
            }
            const uint32_t tileb = sD_u + (uint32_t)buf * 2 * DT_BYTES;
            const uint32_t nbuf = sN_u + (uint32_t)buf * 2 * H_CHUNK;
            float dar[8], daz[8], dan[8], danr[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float dh = acc[i] + pre[i];
                dan[i] = dh * c_n[i];
                dar[i] = dan[i] * c_r[i];
                daz[i] = dh * c_z[i];
                danr[i] = dan[i] * vr[i];
                dhz[i] = dh * vz[i];
                __nv_bfloat16 hi, lo;
                split_bf16(dar[i], hi, lo);
                tc::sts_bf16(tileb + e_off[i], hi);
                tc::sts_bf16(tileb + DT_BYTES + e_off[i], lo);
                split_bf16(daz[i], hi, lo);
                tc::sts_bf16(tileb + H_CHUNK + e_off[i], hi);
                tc::sts_bf16(tileb + DT_BYTES + H_CHUNK + e_off[i], lo);
                split_bf16(danr[i], hi, lo);
                tc::sts_bf16(tileb + 2 * H_CHUNK + e_off[i], hi);
                tc::sts_bf16(tileb + DT_BYTES + 2 * H_CHUNK + e_off[i], lo);
            }
            if (tid == 0) SCANX_TS(9);
            tc::tcgen05_fence_before();
            tc::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(epi_done);
            // the ring slot is released only HERE, after this step's results (which consume every value loaded from the slot)
            // have been written: an arrive right behind the loads was seen to overtake them (the loads sat in the LSU queue behind
            // the previous step's global stores), so the producer's next bulk copy replaced the slot before it had been read
            if (lane == 0) tc::mbar_arrive(&in_empty[s % NSB]);
            if (tid == 0) SCANX_TS(10);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                __nv_bfloat16 hi, lo;
                split_bf16(dan[i], hi, lo);
                tc::sts_bf16(nbuf + e_off[i], hi);
                tc::sts_bf16(nbuf + H_CHUNK + e_off[i], lo);
            }
            tc::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(st_done);
#pragma unroll
            for (int i = 0; i < 8; ++i) { sb_r += dar[i]; sb_z += daz[i]; sb_n += dan[i]; sb_nr += danr[i]; }
        }
        float* dbi = p.db_ih + (int64_t)d * p.dir_stride;
        float* dbh = p.db_hh + (int64_t)d * p.dir_stride;
        atomicAdd(dbi + unit, sb_r); atomicAdd(dbi + H + unit, sb_z); atomicAdd(dbi + 2 * H + unit, sb_n);
        atomicAdd(dbh + unit, sb_r); atomicAdd(dbh + H + unit, sb_z); atomicAdd(dbh + 2 * H + unit, sb_nr);
        if (p.dh0) {
            // gradient of the initial hidden state = z-carry of the last step + W_hh^T dgh of the last step (one more product)
            float acc[8];
            reduce_partials(T, acc);
#pragma unroll
            for (int i = 0; i < 8; ++i) p.dh0[((int64_t)d * B + tile * NB + c0 + i) * H + unit] = dhz[i] + acc[i];
        }
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::cluster_sync_all();
    if (warp == EPI_WARPS) tc::tmem_dealloc(tmem, 512);
}

// ping-pong form: both parts (hi, lo) of a 16-row dgh sub-tile against one row block: 24 MMAs with N = 16
__device__ __forceinline__ void bwd2_issue_block(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_hi, uint64_t desc_lo) {
    constexpr uint32_t idesc = tc::umma_idesc_bf16(128, NBS);
#pragma unroll
    for (int part = 0; part < 2; ++part) {
#pragma unroll
        for (int g = 0; g < 3; ++g) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
                tcs::umma_bf16_ts(tmem_d, tmem_a + (uint32_t)((g * 4 + kk) * 8),
                                  (part ? desc_lo : desc_hi) + (uint64_t)(g * (HS_CHUNK >> 4) + 2 * kk), idesc,
                                  (part == 0 && g == 0 && kk == 0) ? 0u : 1u);
        }
    }
}

template <int H>
__global__ void __launch_bounds__(THREADS, 1) gru_scanx_bwd2_kernel(const __grid_constant__ BwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int CS = H / 64, NKH = H / 128, NRB = 2 * NKH;
    // Ping-pong form: the 32-row tile is worked as two 16-row SUB-TILES.  The thread <-> data mapping of the single-tile kernel
    // already splits by sub-tile - epilogue warps 0-3 touch only batch columns 0-15 (as partial-sum routers AND as owners),
    // warps 4-7 only columns 16-31 - so the two warp groups are simply decoupled: own dgh operand tiles (N = 16), own
    // accumulators, own barriers; the control thread alternates the sub-tiles, and while the tensor pipe multiplies one
    // sub-tile the other group routes / reduces / does its gate-gradient math.
    constexpr int DT_BYTES = 3 * HS_CHUNK;                // one part of one dgh sub-tile
    constexpr int SUBD = 4 * DT_BYTES;                     // [2 buf][2 part] of one sub-tile
    constexpr int SUBN = 4 * HS_CHUNK;
    const int B = p.B, T = p.T;
    uint8_t* sD = smem;                                    // [2 sub][2 buf][2 part][3 gates][HS_CHUNK]
    uint8_t* sN = sD + (size_t)2 * SUBD;                   // [2 sub][2 buf][2 part][HS_CHUNK]   da_n (dgi n-gate rows, store only)
    uint8_t* sR = sN + (size_t)2 * SUBN;                   // [2 buf][4 src][8 cg][64 j] float4 (column groups 0-3: sub-tile 0, 4-7: sub-tile 1)
    uint8_t* sIn = sR + (size_t)2 * RECV_BYTES;            // [NSB][G | YB | dY]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sIn + (size_t)NSB * BWD_STAGE);
    uint64_t* recv_full = bars;        // [2 sub][2 buf]
    uint64_t* mma_a = bars + 4;        // [2 sub] row blocks of the other k half done
    uint64_t* mma_b = bars + 6;        // [2 sub] all row blocks done
    uint64_t* epi_done = bars + 8;     // [2 sub]
    uint64_t* st_done = bars + 10;     // [2 sub]
    uint64_t* in_full = bars + 12;     // [NSB]
    uint64_t* in_empty = bars + 12 + NSB;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12 + 2 * NSB);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t c = tc::cluster_ctarank();
    const int cluster_id = blockIdx.x / CS;
    const int ntiles = B / NB;
    const int d = cluster_id / ntiles, tile = cluster_id % ntiles;
    const bool top = p.dlogits != nullptr;
    const int kh_own = (int)c >> 1;                        // k half that contains this CTA's own units (H = 256)

    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) tc::mbar_init(&recv_full[i], 1);
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&mma_a[i], 1); tc::mbar_init(&mma_b[i], 1);
            tc::mbar_init(&epi_done[i], EPI_WARPS / 2); tc::mbar_init(&st_done[i], EPI_WARPS / 2);
        }
        for (int i = 0; i < NSB; ++i) { tc::mbar_init(&in_full[i], 1); tc::mbar_init(&in_empty[i], EPI_WARPS); }
        tc::fence_mbar_init();
    }
    if (warp == EPI_WARPS) tc::tmem_alloc(tmem_slot, 512);
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::cluster_sync_all();
    tc::tcgen05_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (warp < EPI_WARPS)
        tcs::load_weights_to_tmem(p.WTimg + ((size_t)d * CS + c) * 128 * (NRB * 192), NRB * 192, tmem, A_COL, warp, lane);
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();

    if (warp == EPI_WARPS + 1) {
        if (tc::elect_one()) {
            bool ok = true;
            for (int s = 0; s < T; ++s) {
                const int st = s % NSB;
                if (s >= NSB && ok) ok = tc::mbar_wait(&in_empty[st], ((s / NSB) - 1) & 1, p.dbg, 0x2300 + (s & 0xff));
                const int t = d == 0 ? T - 1 - s : s;
                const bool first = d == 0 ? t == 0 : t == T - 1;
                uint8_t* dst = sIn + (size_t)st * BWD_STAGE;
                tc::mbar_arrive_expect_tx(&in_full[st], (uint32_t)(G_BLOCK + (top ? 0 : DY_BLOCK) + (first ? 0 : YB_BLOCK)));
                const size_t blk = blk_index(d, tile, t, (int)c, ntiles, T, CS);
                tc::bulk_g2s(dst, reinterpret_cast<const uint8_t*>(p.GX) + blk * G_BLOCK, G_BLOCK, &in_full[st]);
                if (!top) tc::bulk_g2s(dst + G_BLOCK + YB_BLOCK, reinterpret_cast<const uint8_t*>(p.dYBX) + blk * DY_BLOCK, DY_BLOCK, &in_full[st]);
                if (!first) {
                    const size_t pblk = blk_index(d, tile, d == 0 ? t - 1 : t + 1, (int)c, ntiles, T, CS);
                    tc::bulk_g2s(dst + G_BLOCK, reinterpret_cast<const uint8_t*>(p.YBX) + pblk * YB_BLOCK, YB_BLOCK, &in_full[st]);
                }
            }
        }
    } else if (warp == EPI_WARPS) {
        if (tc::elect_one()) {
            bool ok = true;
            auto store_tile = [&](int sub, int step) {        // 16-row boxes (the tensor maps of this form have box 64 x 16)
                const int tt = d == 0 ? T - 1 - step : step;
                const int row = tt * B + tile * NB + sub * NBS;
                const uint8_t* tb = sD + (size_t)sub * SUBD + (size_t)(step & 1) * 2 * DT_BYTES;
                const uint8_t* nb = sN + (size_t)sub * SUBN + (size_t)(step & 1) * 2 * HS_CHUNK;
                const int cu = (int)c * UNITS;
                tc::tma_store_2d(&p.tmGIh, tb, d * 3 * H + cu, row);                               // da_r
                tc::tma_store_2d(&p.tmGIl, tb + DT_BYTES, d * 3 * H + cu, row);
                tc::tma_store_2d(&p.tmGIh, tb + HS_CHUNK, d * 3 * H + H + cu, row);                // da_z
                tc::tma_store_2d(&p.tmGIl, tb + DT_BYTES + HS_CHUNK, d * 3 * H + H + cu, row);
                tc::tma_store_2d(&p.tmGIh, nb, d * 3 * H + 2 * H + cu, row);                       // da_n
                tc::tma_store_2d(&p.tmGIl, nb + HS_CHUNK, d * 3 * H + 2 * H + cu, row);
                tc::tma_store_2d(&p.tmGNh, tb + 2 * HS_CHUNK, d * H + cu, row);                    // da_n * r
                tc::tma_store_2d(&p.tmGNl, tb + DT_BYTES + 2 * HS_CHUNK, d * H + cu, row);
                tc::tma_store_commit();
            };
            const uint32_t db0 = tc::smem_u32(sD);
            const int Tend = T + (p.dh0 ? 1 : 0);          // one more product (no gate math) when d(h0) is wanted
            for (int s = 1; s < Tend; ++s) {
                const int pb = (s - 1) & 1;
#pragma unroll
                for (int sub = 0; sub < 2; ++sub) {
                    if (ok) ok = tc::mbar_wait(&epi_done[sub], (s - 1) & 1, p.dbg, 0x2700 + (s & 0xff));
                    if (sub == 0) SCANX_TS(0);
                    tc::tcgen05_fence_after();
                    tc::mbar_arrive_expect_tx(&recv_full[sub * 2 + (s & 1)], (uint32_t)(CS - 1) * 4096u);
                    const uint32_t tb = db0 + (uint32_t)sub * SUBD + (uint32_t)pb * 2 * DT_BYTES;
                    const uint64_t dhi = tc::umma_desc_k_sw128(tb), dlo = tc::umma_desc_k_sw128(tb + DT_BYTES);
                    const uint32_t td = tmem + (uint32_t)(sub * 4 * NBS);
                    const int ko = 1 - kh_own;
                    bwd2_issue_block(td + (uint32_t)((0 * NKH + ko) * NBS), tmem + A_COL + (uint32_t)((0 * NKH + ko) * 96), dhi, dlo);
                    bwd2_issue_block(td + (uint32_t)((1 * NKH + ko) * NBS), tmem + A_COL + (uint32_t)((1 * NKH + ko) * 96), dhi, dlo);
                    tc::umma_commit(&mma_a[sub]);
                    if (sub == 0) SCANX_TS(1);
                    bwd2_issue_block(td + (uint32_t)((0 * NKH + kh_own) * NBS), tmem + A_COL + (uint32_t)((0 * NKH + kh_own) * 96), dhi, dlo);
                    bwd2_issue_block(td + (uint32_t)((1 * NKH + kh_own) * NBS), tmem + A_COL + (uint32_t)((1 * NKH + kh_own) * 96), dhi, dlo);
                    tma_store_wait_read1();            // the tile of THIS sub-tile stored two steps ago has been read (the other's may be in flight)
                    tc::umma_commit(&mma_b[sub]);
                    if (sub == 0) SCANX_TS(3);
                    if (ok) ok = tc::mbar_wait(&st_done[sub], (s - 1) & 1, p.dbg, 0x2a00 + (s & 0xff));
                    store_tile(sub, s - 1);
                }
            }
            if (Tend == T) {
                for (int sub = 0; sub < 2; ++sub) {
                    if (ok) ok = tc::mbar_wait(&epi_done[sub], (T - 1) & 1, p.dbg, 0x2700);
                    if (ok) ok = tc::mbar_wait(&st_done[sub], (T - 1) & 1, p.dbg, 0x2a00);
                    store_tile(sub, T - 1);
                }
            }
            tc::tma_store_wait_all();
        }
    } else {
        // ---- epilogue.  Owner role: unit j = (warp & 1)*32 + lane of this CTA, batch columns [8*(warp >> 1), +8).
        //      Partial-sum role: TMEM lane quarter q = warp & 3 -> output unit k = 128*kh + 32q + lane, columns [16*half, +16).
        const int q = warp & 3, half = warp >> 2;          // half == sub-tile of this warp (router AND owner roles)
        const int sub = half;
        const int j = (warp & 1) * 32 + lane;
        const int unit = (int)c * UNITS + j;
        const int c0 = 8 * (warp >> 1);
        const int tid = threadIdx.x;
        float dhz[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) dhz[i] = 0.f;
        float h_avg[8], h_max[8];
        int h_arg[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { h_avg[i] = 0.f; h_max[i] = 0.f; h_arg[i] = -1; }
        if (top) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int b = tile * NB + c0 + i;
                float dl = 0.f, dm = 0.f, da = 0.f;
                for (int cc = 0; cc < p.C; ++cc) {
                    const float g = p.dlogits[(int64_t)b * p.C + cc];
                    const float* w = p.lin_w + (int64_t)cc * 3 * H;
                    dl = fmaf(g, w[unit], dl); dm = fmaf(g, w[H + unit], dm); da = fmaf(g, w[2 * H + unit], da);
                }
                dhz[i] = dl;
                h_avg[i] = da / (float)T; h_max[i] = dm; h_arg[i] = p.arg[(int64_t)b * H + unit];
            }
        }
        float sb_r = 0.f, sb_z = 0.f, sb_n = 0.f, sb_nr = 0.f;
        uint32_t e_off[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) e_off[i] = tc::sw128_offset(c0 - 16 * sub + i, j);      // row inside the 16-row sub-tile
        // partial-sum destination inside a receive buffer: [src = c][cg = 4*half + i][jd] float4, jd = (q & 1)*32 + lane
        const uint32_t r_off = (((uint32_t)c * 8 + 4 * half) * 64 + (uint32_t)((q & 1) * 32 + lane)) * 16;
        const uint32_t sIn_u = tc::smem_u32(sIn), sR_u = tc::smem_u32(sR), sD_u = tc::smem_u32(sD), sN_u = tc::smem_u32(sN);
        bool ok = true;
        // recurrent part of dh for this thread's (unit, 8 columns) at step s: every CTA's row blocks -> partial sums (hi + lo
        // blocks of one k share a lane) -> owner's receive buffer (st.async / local store) -> sum over the CS sources
        auto reduce_partials = [&](int s, float (&acc)[8]) {
            const int buf = s & 1;
            const uint32_t rb_local = sR_u + (uint32_t)buf * RECV_BYTES;
            const uint32_t rbar_l = tc::smem_u32(&recv_full[sub * 2 + buf]);
            auto route = [&](int kh) {
                float vh[16], vl[16];
                const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(sub * 4 * NBS);
                tmem_ld16f(ta + (uint32_t)((0 * NKH + kh) * NBS), vh);
                tmem_ld16f(ta + (uint32_t)((1 * NKH + kh) * NBS), vl);
                tmem_ld_wait_pin(vh, vl);
                const uint32_t dest = (uint32_t)(2 * kh + (q >> 1));
                const uint32_t lp = rb_local + r_off;
                if (dest == c) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        tc::sts_f4(lp + (uint32_t)(i * 64 * 16),
                                   make_float4(vh[4 * i] + vl[4 * i], vh[4 * i + 1] + vl[4 * i + 1], vh[4 * i + 2] + vl[4 * i + 2], vh[4 * i + 3] + vl[4 * i + 3]));
                } else {
                    const uint32_t ra = tc::mapa_u32(lp, dest), rb = tc::mapa_u32(rbar_l, dest);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        uint4 u;
                        u.x = __float_as_uint(vh[4 * i] + vl[4 * i]); u.y = __float_as_uint(vh[4 * i + 1] + vl[4 * i + 1]);
                        u.z = __float_as_uint(vh[4 * i + 2] + vl[4 * i + 2]); u.w = __float_as_uint(vh[4 * i + 3] + vl[4 * i + 3]);
                        tc::st_async_v4(ra + (uint32_t)(i * 64 * 16), u, rb);
                    }
                }
            };
            if (NKH == 2) {
                if (tid == 0) SCANX_TS(4);
                if (ok) ok = tc::mbar_wait(&mma_a[sub], (s - 1) & 1, p.dbg, 0x2800 + (s & 0xff));
                tc::tcgen05_fence_after();
                route(1 - kh_own);
            }
            if (ok) ok = tc::mbar_wait(&mma_b[sub], (s - 1) & 1, p.dbg, 0x2900 + (s & 0xff));
            if (tid == 0) SCANX_TS(7);
            tc::tcgen05_fence_after();
            route(NKH == 2 ? kh_own : 0);
            tc::tcgen05_fence_before();
            asm volatile("bar.sync %0, 128;" ::"r"(1 + sub) : "memory");      // this CTA's own contributions (of this warp group) are in the buffer
            if (ok) ok = tc::mbar_wait_cluster(&recv_full[sub * 2 + buf], ((s - 1) >> 1) & 1, p.dbg, 0x2b00 + (s & 0xff));
            if (tid == 0) SCANX_TS(8);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
            for (int src = 0; src < CS; ++src) {
                const uint32_t rp = rb_local + (uint32_t)((((src * 8 + 2 * (warp >> 1)) * 64) + j) * 16);
                const float4 x0 = tc::lds_f4(rp), x1 = tc::lds_f4(rp + 64 * 16);
                acc[0] += x0.x; acc[1] += x0.y; acc[2] += x0.z; acc[3] += x0.w; acc[4] += x1.x; acc[5] += x1.y; acc[6] += x1.z; acc[7] += x1.w;
            }
        };
        for (int s = 0; s < T; ++s) {
            const int t = d == 0 ? T - 1 - s : s;
            const bool first = d == 0 ? t == 0 : t == T - 1;
            float vr[8], vz[8], vn[8], vhn[8], vhp[8], vdy[8];
            {
                const int st = s % NSB;
                if (ok) ok = tc::mbar_wait(&in_full[st], (s / NSB) & 1, p.dbg, 0x2200 + (s & 0xff));
                const uint32_t gp = sIn_u + (uint32_t)st * BWD_STAGE + 32u * tid;
                const float4 a0 = tc::lds_f4(gp), a1 = tc::lds_f4(gp + 16), b0 = tc::lds_f4(gp + 8192), b1 = tc::lds_f4(gp + 8192 + 16),
                             n0 = tc::lds_f4(gp + 16384), n1 = tc::lds_f4(gp + 16384 + 16), m0 = tc::lds_f4(gp + 24576), m1 = tc::lds_f4(gp + 24576 + 16);
                float4 p0 = make_float4(0.f, 0.f, 0.f, 0.f), p1 = p0, y0 = p0, y1 = p0;
                if (!first) { p0 = tc::lds_f4(gp + G_BLOCK); p1 = tc::lds_f4(gp + G_BLOCK + 16); }
                if (!top) { y0 = tc::lds_f4(gp + G_BLOCK + YB_BLOCK); y1 = tc::lds_f4(gp + G_BLOCK + YB_BLOCK + 16); }
                vr[0] = a0.x; vr[1] = a0.y; vr[2] = a0.z; vr[3] = a0.w; vr[4] = a1.x; vr[5] = a1.y; vr[6] = a1.z; vr[7] = a1.w;
                vz[0] = b0.x; vz[1] = b0.y; vz[2] = b0.z; vz[3] = b0.w; vz[4] = b1.x; vz[5] = b1.y; vz[6] = b1.z; vz[7] = b1.w;
                vn[0] = n0.x; vn[1] = n0.y; vn[2] = n0.z; vn[3] = n0.w; vn[4] = n1.x; vn[5] = n1.y; vn[6] = n1.z; vn[7] = n1.w;
                vhn[0] = m0.x; vhn[1] = m0.y; vhn[2] = m0.z; vhn[3] = m0.w; vhn[4] = m1.x; vhn[5] = m1.y; vhn[6] = m1.z; vhn[7] = m1.w;
                vhp[0] = p0.x; vhp[1] = p0.y; vhp[2] = p0.z; vhp[3] = p0.w; vhp[4] = p1.x; vhp[5] = p1.y; vhp[6] = p1.z; vhp[7] = p1.w;
                vdy[0] = y0.x; vdy[1] = y0.y; vdy[2] = y0.z; vdy[3] = y0.w; vdy[4] = y1.x; vdy[5] = y1.y; vdy[6] = y1.z; vdy[7] = y1.w;
                if (first && p.h0) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) vhp[i] = p.h0[((int64_t)d * B + tile * NB + c0 + i) * H + unit];
                }
                if (top) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) vdy[i] = h_avg[i] + (h_arg[i] == t ? h_max[i] : 0.f);
                }
            }
            float c_n[8], c_r[8], c_z[8], pre[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float r = vr[i], z = vz[i], n = vn[i];
                c_n[i] = (1.f - z) * (1.f - n * n);
                c_r[i] = vhn[i] * r * (1.f - r);
                c_z[i] = (vhp[i] - n) * z * (1.f - z);
                pre[i] = dhz[i] + vdy[i];
            }
            float acc[8];
            const int buf = s & 1;
            if (s > 0) reduce_partials(s, acc);
            else {
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] = 0.f;
            }
            const uint32_t tileb = sD_u + (uint32_t)sub * SUBD + (uint32_t)buf * 2 * DT_BYTES;
            const uint32_t nbuf = sN_u + (uint32_t)sub * SUBN + (uint32_t)buf * 2 * HS_CHUNK;
            float dar[8], daz[8], dan[8], danr[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float dh = acc[i] + pre[i];
                dan[i] = dh * c_n[i];
                dar[i] = dan[i] * c_r[i];
                daz[i] = dh * c_z[i];
                danr[i] = dan[i] * vr[i];
                dhz[i] = dh * vz[i];
                __nv_bfloat16 hi, lo;
                split_bf16(dar[i], hi, lo);
                tc::sts_bf16(tileb + e_off[i], hi);
                tc::sts_bf16(tileb + DT_BYTES + e_off[i], lo);
                split_bf16(daz[i], hi, lo);
                tc::sts_bf16(tileb + HS_CHUNK + e_off[i], hi);
                tc::sts_bf16(tileb + DT_BYTES + HS_CHUNK + e_off[i], lo);
                split_bf16(danr[i], hi, lo);
                tc::sts_bf16(tileb + 2 * HS_CHUNK + e_off[i], hi);
                tc::sts_bf16(tileb + DT_BYTES + 2 * HS_CHUNK + e_off[i], lo);
            }
            if (tid == 0) SCANX_TS(9);
            tc::tcgen05_fence_before();
            tc::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&epi_done[sub]);
            // the ring slot is released only HERE, after this step's results (which consume every value loaded from the slot)
            // have been written: an arrive right behind the loads was seen to overtake them (the loads sat in the LSU queue behind
            // the previous step's global stores), so the producer's next bulk copy replaced the slot before it had been read
            if (lane == 0) tc::mbar_arrive(&in_empty[s % NSB]);
            if (tid == 0) SCANX_TS(10);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                __nv_bfloat16 hi, lo;
                split_bf16(dan[i], hi, lo);
                tc::sts_bf16(nbuf + e_off[i], hi);
                tc::sts_bf16(nbuf + HS_CHUNK + e_off[i], lo);
            }
            tc::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&st_done[sub]);
#pragma unroll
            for (int i = 0; i < 8; ++i) { sb_r += dar[i]; sb_z += daz[i]; sb_n += dan[i]; sb_nr += danr[i]; }
        }
        float* dbi = p.db_ih + (int64_t)d * p.dir_stride;
        float* dbh = p.db_hh + (int64_t)d * p.dir_stride;
        atomicAdd(dbi + unit, sb_r); atomicAdd(dbi + H + unit, sb_z); atomicAdd(dbi + 2 * H + unit, sb_n);
        atomicAdd(dbh + unit, sb_r); atomicAdd(dbh + H + unit, sb_z); atomicAdd(dbh + 2 * H + unit, sb_nr);
        if (p.dh0) {
            // gradient of the initial hidden state = z-carry of the last step + W_hh^T dgh of the last step (one more product)
            float acc[8];
            reduce_partials(T, acc);
#pragma unroll
            for (int i = 0; i < 8; ++i) p.dh0[((int64_t)d * B + tile * NB + c0 + i) * H + unit] = dhz[i] + acc[i];
        }
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::cluster_sync_all();
    if (warp == EPI_WARPS) tc::tmem_dealloc(tmem, 512);
}

static inline cudaError_t launch_bwd(const BwdParams& p_in, cudaStream_t st) {
    BwdParams p = p_in;
    if ((p.H != 128 && p.H != 256) || p.B % NB != 0) return cudaErrorInvalidValue;
    // H = 256: the ping-pong form (two 16-row sub-tiles); BIGRU_X3_BWD=single selects the single-tile kernel
    static const bool single = [] { const char* e = getenv("BIGRU_X3_BWD"); return e && e[0] == 's'; }();
    const bool pp = p.H == 256 && !single;
    {
        const uint32_t box[2] = {64u, (uint32_t)(pp ? NBS : NB)};
        const uint64_t d1[2] = {(uint64_t)p.D * 3 * p.H, (uint64_t)p.T * p.B};
        const uint64_t s1[1] = {(uint64_t)p.D * 3 * p.H * 2};
        const uint64_t d2[2] = {(uint64_t)p.D * p.H, (uint64_t)p.T * p.B};
        const uint64_t s2[1] = {(uint64_t)p.D * p.H * 2};
        if (make_tmap_bf16(&p.tmGIh, p.dgi_hi, 2, d1, s1, box) != 0 || make_tmap_bf16(&p.tmGIl, p.dgi_lo, 2, d1, s1, box) != 0 ||
            make_tmap_bf16(&p.tmGNh, p.dghn_hi, 2, d2, s2, box) != 0 || make_tmap_bf16(&p.tmGNl, p.dghn_lo, 2, d2, s2, box) != 0)
            return cudaErrorInvalidValue;
    }
    const int CS = p.H / UNITS;
    const size_t smem = bwd_smem_bytes();
    void (*kern)(BwdParams) = p.H == 128 ? gru_scanx_bwd_kernel<128> : (pp ? gru_scanx_bwd2_kernel<256> : gru_scanx_bwd_kernel<256>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(p.D * (p.B / NB) * CS));
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, p);
}

// ---- weight images ------------------------------------------------------------------------------
// forward: img[((d*CS + c)*128 + row)*3H + g*H + k] = split_{row/64}(W_hh[g*H + 64c + row%64][k])
// backward: img[((d*CS + c)*128 + i)*(NRB*192) + rb*192 + g*64 + jj] = split_{rb/NKH}(W_hh[g*H + 64c + jj][128*(rb%NKH) + i])
__global__ void pack_whh_images_kernel(const float* __restrict__ w_hh, __nv_bfloat16* __restrict__ fimg, __nv_bfloat16* __restrict__ bimg, int H) {
    const int CS = H / 64, NKH = H / 128, NRB = 2 * NKH;
    const int64_t per_cta = (int64_t)128 * 3 * H;
    const int64_t total = (int64_t)CS * per_cta;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i / per_cta);
        const int64_t r = i % per_cta;
        {   // forward image element
            const int row = (int)(r / (3 * H)), col = (int)(r % (3 * H));
            const int g = col / H, k = col % H, part = row >> 6, jj = row & 63;
            const float w = w_hh[((int64_t)g * H + 64 * c + jj) * H + k];
            __nv_bfloat16 hi, lo;
            split_bf16(w, hi, lo);
            fimg[i] = part ? lo : hi;
        }
        {   // backward image element (NRB*192 == 3H elements per lane as well)
            const int lane_i = (int)(r / (NRB * 192)), col = (int)(r % (NRB * 192));
            const int rb = col / 192, kq = col % 192, g = kq / 64, jj = kq % 64;
            const int part = rb / NKH, kh = rb % NKH;
            const float w = w_hh[((int64_t)g * H + 64 * c + jj) * H + 128 * kh + lane_i];
            __nv_bfloat16 hi, lo;
            split_bf16(w, hi, lo);
            bimg[i] = part ? lo : hi;
        }
    }
}

}  // namespace tcx
