// tools/experimental/tc_scanx_pingpong.cuh - NOT part of the product (not included by libbigru_b200.so).
// Ping-pong form of the x3 forward scan (two 16-row sub-tiles alternating on the tensor pipe / epilogue warps / DSMEM, the
// lo-row warps doubling as st.async senders).  Measured on B200: 3.12 us/step against 3.48 for the single-tile kernel of
// csrc/tc_scan_x.cuh (the step stays DSMEM-bound: every CTA ships 24 KB of h per step at ~12-16 B/clk), bit-for-bit correct
// in the stand-alone test and in 90 stress runs, but at the full C1 shape inside the train step the logits were off by
// 5e-4..8e-4 (run-to-run different: a race that only shows with 32 clusters resident) - not adopted.  Kept as the record of
// the experiment; it compiles when pasted back before launch_fwd in tc_scan_x.cuh (namespace tcx).
#if 0

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
                        __syncwarp();
                        if (lane == 0) tc::mbar_arrive(&in_empty[sub * NS2 + st]);
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


#endif
