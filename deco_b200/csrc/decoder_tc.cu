// Per-pixel AdaLN-MLP pixel decoder on the 5th-generation tensor cores, with the CFG-batched sampler update fused into
// its epilogue.
//
// Replaces (reference, /root/reference):
//   src/models/transformer/dit_c2i_DeCo.py:212-248   NerfEmbedder (constant positional table)
//   src/models/transformer/dit_c2i_DeCo.py:395-415   SimpleMLPAdaLN.forward (input_proj, res blocks, final layer)
//   src/models/transformer/dit_c2i_DeCo.py:313-317   ResBlock.forward (LayerNorm, adaLN modulate, MLP, gated residual)
//   src/models/transformer/dit_c2i_DeCo.py:329-332   decoder FinalLayer;  :501-509 reshape / transpose / F.fold
//   src/diffusion/base/guidance.py:3-6 + src/diffusion/flow_matching/sampling.py:89-104 (+ adam_sampling.py:104-117,
//   src/models/autoencoder/base.py:32-34) when the sampler step is fused: pred = u + g (c - u), v = c0 pred + c1 p1,
//   x += dt v, optional fp2uint8 -- BASELINE north_star (4): "the sampler step, fused into the decoder epilogue".
//
// Geometry.  A TILE is 128 pixels = 8 rows x 16 columns of one 16 x 16 patch (half a token); pixel <-> TMEM lane <->
// epilogue thread, so LayerNorm over the 32 channels, the gated residual and the sampler update are thread-local and the
// residual x[32] lives in fp32 registers for the whole chain.  Every matrix product of the chain is a tcgen05.mma with
// M = 128 pixels, K = 32 (or 16), N = 16 / 32 / 64:
//   * the A operand is either the tile of silu(cond_embed(s)) (written by the cond_embed GEMM's SiLU epilogue, fetched by
//     TMA straight into the 32-byte-swizzled K-major operand layout: no thread touches it) or the activation the
//     epilogue threads just produced, packed to bf16 and stored into TENSOR memory (tcgen05.st; TS-form MMA);
//   * the B operands (all weights, 14 KB per res block) stay resident in shared memory;
//   * biases ride on a K = 16 "bias MMA": A = a constant tile with ones in columns 0 and 1, B = (bf16 hi, bf16 lo) of the
//     bias, so an accumulator never needs an fp32 bias pass.
// Algebra done once on the host (deco_b200/denoiser.py::pack_decoder_tc), per res block with adaLN rows (shift, scale,
// gate), LayerNorm affine (g, b), MLP (W0, b0, W2, b2) and a = silu(cond):
//   sc' = 1 + scale = Wscale a + (1 + bscale)                                   -> accumulator SC
//   gt  = Wgate a + bgate                                                       -> accumulator GT
//   W0 h + b0  with  h = (LN(x) g + b)(1 + scale) + shift
//           = (W0 diag(g)) (LN(x) . sc') + (W0 diag(b) Wscale + W0 Wshift) a + (W0 (b (1 + bscale)) + W0 bshift + b0)
//     i.e. ONE accumulator H fed by a TS MMA (the normalised, scaled activation), an SS MMA (composite weight against the
//     condition tile) and the bias MMA; the three are pre-multiplied by 1/2 so that silu(v) = hv + hv tanh(hv), hv = v/2,
//     costs one FMA and one MUFU per channel.
//   NerfEmbedder + input_proj collapse into x = T'[pixel] + W' rgb with T' = Win (Wpos table + bx) + bin (fp32 table,
//     [256][32]) and W' = Win Wrgb ([32][3]).
// Per tile the epilogue threads do: LN statistics, 2 multiplies per channel for the modulated activation, one FMA + MUFU
// per channel for SiLU, one FMA per channel for the gated residual -- ~7 FP32 operations per channel and res block.
//
// Pipeline.  Persistent kernel, one CTA per SM, 20 warps: 4 SLOTS of 4 epilogue warps (a slot works on one tile at a time;
// its 128 TMEM columns hold SC | GT | H | the bf16 A operand) and one MMA issuer warp PER SLOT, which also fetches the
// slot's condition tiles by TMA (double buffered, two tiles ahead).  A tile is 2R + 1 MMA <-> epilogue round trips (R res blocks); the four
// slots interleave so that one slot's round trip latency hides behind the other three slots' arithmetic.  One thread can
// issue a tcgen05.mma only every ~58 clocks whatever its size (scripts/tmem_bench.cu: N = 32 runs at its 16-clock floor
// only when four warps issue), hence one issuer per slot; each issuer commits the MMAs the waiting epilogue needs FIRST
// and then pre-issues whatever no longer depends on the epilogue (next block's scale MMA as soon as the current one has
// been read, next tile's scale / gate MMAs after the final layer), with descriptors precomputed outside the loop.
//
// Sampler fusion (mode "pair"): rows b (uncond) and b + B (cond) of the CFG batch share the image x; a slot runs the two
// tiles back to back, keeps the first result in 3 registers and applies guidance + the multistep update to the fp32
// state in place; the bf16 network output never exists in HBM.
#include "tcgen05.cuh"
#include "tma_host.cuh"

namespace deco {
namespace dtc {

constexpr int kSlots = 4;
constexpr int kEpiWarps = 16;
constexpr int kThreads = (kEpiWarps + kSlots) * 32;  // + one MMA issuer per slot (which also fetches its slot's condition tiles)
// (20 warps: registers are allocated to warps in groups of four, a 21st warp would cost every thread 16 registers)
constexpr int kMaxR = 6;
constexpr int kHx = 32;

// tensor-memory columns of a slot
constexpr uint32_t kColSC = 0, kColGT = 32, kColH = 64, kColA = 96, kSlotCols = 128;

// weight image (bytes; host-packed, every tile 256-byte aligned, 32-byte-swizzled K-major: sw32_offset)
constexpr uint32_t kOnes = 0;                         // A [128 x 16]: ones in columns 0, 1
constexpr uint32_t kBlock0 = 4096;
constexpr uint32_t kWsg = 0;                          // B [64 x 32]: rows 0-31 scale, 32-63 gate
constexpr uint32_t kW0 = 4096;                        // B [32 x 32]: 1/2 W0 diag(g)
constexpr uint32_t kC0 = 6144;                        // B [32 x 32]: 1/2 (W0 diag(b) Wscale + W0 Wshift)
constexpr uint32_t kW2 = 8192;                        // B [32 x 32]
constexpr uint32_t kBsg = 10240;                      // B [64 x 16]: (hi, lo) of 1 + bscale | bgate
constexpr uint32_t kB0 = 12288;                       // B [32 x 16]
constexpr uint32_t kB2 = 13312;                       // B [32 x 16]
constexpr uint32_t kBlockBytes = 14336;
constexpr uint32_t kWf = 0;                           // B [16 x 32] final linear (rows 3-15 zero)
constexpr uint32_t kBf = 1024;                        // B [16 x 16]
constexpr uint32_t kFinalBytes = 2048;
constexpr int kTabPitch = 36;                         // floats per row of T' (bank-conflict-free 128-bit reads)
constexpr uint32_t kTabBytes = 256 * kTabPitch * 4;   // T' [256][36]
constexpr uint32_t kWrgbBytes = 32 * 4 * 4;           // W' [32][4]
constexpr uint32_t kYTile = 128 * 32 * 2;             // one condition tile: [128 pixels x 32 channels] bf16

__host__ __device__ inline uint32_t weight_bytes(int R) { return kBlock0 + (uint32_t)R * kBlockBytes + kFinalBytes; }
__host__ __device__ inline uint32_t blob_bytes(int R) { return weight_bytes(R) + kTabBytes + kWrgbBytes; }
__host__ __device__ inline uint32_t smem_bytes(int R) { return blob_bytes(R) + kSlots * 2 * kYTile + 512 /*barriers*/ + 1024 /*align*/; }

struct Params {
    const float* x;              // [B, 3, H, W] fp32 image state (pair mode) or [rows, 3, H, W] network input
    const void* blob;
    void* out;                   // plain mode: [rows, 3, H, W] bf16 / fp32
    int out_bf16;
    int R, H, W, Hp, Wp;
    long long tokens;            // rows * L
    // ---- pair mode (sampler step fused): rows = 2B, row b = uncond, row b + B = cond
    int pair;
    long long tokens_half;       // B * L
    const float* dev;            // device {g, dt, c0, c1} or null (then the host scalars below)
    float g, dt, c0, c1;
    const float* x_base;         // state the update is applied to, or null (= x).  Heun corrector: the net sees x_hat, the
                                 // update x + dt (v + v_hat) / 2 starts from x (sampling.py:289-291)
    const float* p1;             // previous prediction (Adams order 2, Heun predictor velocity) or null
    float* x_out;                // new state (may alias x)
    float* pred_out;             // optional: guided prediction (may alias p1)
    uint8_t* u8_out;             // optional: fp2uint8(new state)
};

__device__ __forceinline__ float tanh_approx(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint8_t to_u8(float v) {
    // fp2uint8 (src/models/autoencoder/base.py:32-34): clamp((x + 1) * 127.5 + 0.5, 0, 255) -> uint8 (truncation)
    float y = fminf(fmaxf(fmaf(v + 1.0f, 127.5f, 0.5f), 0.0f), 255.0f);
    return (uint8_t)y;
}

#ifdef DECO_DTC_TRACE
// clock stamps of CTA 0 / slot 0: region 0 = epilogue thread 0, region 1 = the slot's MMA issuer (scripts/dtc_trace.py)
__device__ long long g_dtc_trace[2 * 4096];
#define DTC_TRACE(region, code)                                                                   \
    do {                                                                                          \
        if (blockIdx.x == 0 && trace_on && trace_n < 2047) {                                      \
            g_dtc_trace[(region) * 4096 + 2 * trace_n] = clock64();                               \
            g_dtc_trace[(region) * 4096 + 2 * trace_n + 1] = (code);                              \
            ++trace_n;                                                                            \
        }                                                                                         \
    } while (0)
#else
#define DTC_TRACE(region, code) do {} while (0)
#endif

__global__ void __launch_bounds__(kThreads, 1)
pixel_decoder_tc_kernel(const __grid_constant__ CUtensorMap ymap, const Params P)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    const int R = P.R;
    const uint32_t sW = base;                                   // weight image
    const uint32_t oTab = weight_bytes(R);
    const float* sTab = reinterpret_cast<const float*>(gen + oTab);
    const float* sWrgb = reinterpret_cast<const float*>(gen + oTab + kTabBytes);
    const uint32_t sY = base + blob_bytes(R);                   // [slot][buf] condition tiles
    const uint32_t bars = sY + kSlots * 2 * kYTile;
    auto d_bar = [&](int k) { return bars + 8u * k; };                          // MMA group complete -> slot k's threads
    auto a_bar = [&](int k) { return bars + 32u + 8u * k; };                    // slot k's A operand / accumulators released
    auto y_full = [&](int k, int b) { return bars + 64u + 8u * (2 * k + b); };
    auto y_empty = [&](int k, int b) { return bars + 128u + 8u * (2 * k + b); };
    auto p_bar = [&](int k) { return bars + 192u + 8u * k; };                   // scale' | gate of a tile's first block ready
    const uint32_t tslot = bars + 224u;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nstage = 2 * R + 1;                               // epilogue -> issuer hand-overs per tile

    // work: items (token, half) dealt to CTAs round-robin, then to the CTA's slots round-robin; pair mode: two tiles per item.
    // All of it is 32-bit arithmetic (tokens * 256 < 2^31 is checked on the host): 64-bit divisions cost hundreds of clocks.
    const int item_tokens = (int)(P.pair ? P.tokens_half : P.tokens);
    const int nitems = item_tokens * 2;
    const int first = blockIdx.x, step = gridDim.x;
    const int nlocal = first < nitems ? (nitems - first + step - 1) / step : 0;
    const int tpi = P.pair ? 2 : 1;                             // tiles per item
    auto slot_tiles = [&](int k) -> int { return nlocal > k ? ((nlocal - k + kSlots - 1) / kSlots) * tpi : 0; };
    // tile ts of slot k -> token row m (in ycond / out rows) and half
    auto tile_of = [&](int k, int ts, int& m, int& half) {
        const int q = (P.pair ? (ts >> 1) : ts) * kSlots + k;   // CTA-local item
        const int it = first + q * step;
        half = it & 1;
        m = (it >> 1) + ((P.pair && (ts & 1)) ? item_tokens : 0);
    };

    if (tid == 0) {
        for (int k = 0; k < kSlots; ++k) {
            mbar_init(d_bar(k), 1);
            mbar_init(p_bar(k), 1);
            mbar_init(a_bar(k), 4);
            for (int b = 0; b < 2; ++b) { mbar_init(y_full(k, b), 1); mbar_init(y_empty(k, b), 1); }
        }
        fence_barrier_init();
        tma_prefetch_desc(&ymap);
    }
    {   // weights + tables: constants, not a predecessor's output -> staged before the dependency wait
        const uint4* src = reinterpret_cast<const uint4*>(P.blob);
        uint4* dst = reinterpret_cast<uint4*>(gen);
        const int n16 = (int)(blob_bytes(R) / 16);
        for (int i = tid; i < n16; i += kThreads) dst[i] = __ldg(src + i);
    }
    if (warp == kEpiWarps) tmem_alloc(tslot, 512);
    fence_proxy_async();          // the generic-proxy weight stores above are read by the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_wait();                   // everything above overlapped the cond_embed GEMM's tail; ycond / x are visible from here
    pdl_launch_dependents();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + (tslot - base));

    if (warp >= kEpiWarps) {
        // ================================================================== MMA issuer of slot k (one elected lane issues)
        const int k = warp - kEpiWarps;
        const int nt = slot_tiles(k);
#ifdef DECO_DTC_TRACE
        int trace_n = 0;
        const bool trace_on = k == 0 && lane == 0;
#endif
        constexpr uint32_t id16 = make_idesc_major(128, 16, 0, 0), id32 = make_idesc_major(128, 32, 0, 0),
                           id64 = make_idesc_major(128, 64, 0, 0);
        auto desc = [&](uint32_t addr) { return make_umma_desc(addr, 16, 256, 6); };
        // descriptors are affine in the shared-memory address: one base per operand kind, the rest is integer offsets in
        // 16-byte units (the address field of the descriptor)
        const uint64_t dW = desc(sW);                                       // weight image base
        const uint64_t dY0 = desc(sY + (uint32_t)(2 * k) * kYTile);         // condition tile, buffer 0 (buffer 1: + kYTile)
        auto wd = [&](uint32_t off) { return dW + (uint64_t)(off >> 4); };
        const uint64_t dOnes = wd(kOnes);
        const uint32_t tcol = tmem + (uint32_t)k * kSlotCols;
        const uint32_t acol = tcol + kColA;
        // D (+)= Y[128 x 32] (2 chunks of 4096 B) . B^T with B = rows [row0, row0 + N) of a K-major tile of `brows` rows per
        // chunk at weight offset boff
        auto mma_y = [&](uint32_t dcol, uint32_t idesc, uint64_t dy, uint32_t boff, uint32_t brows, uint32_t row0, uint32_t acc) {
            umma_bf16(dcol, dy, wd(boff + row0 * 32), idesc, acc);
            umma_bf16(dcol, dy + ((128 * 32) >> 4), wd(boff + brows * 32 + row0 * 32), idesc, 1u);
        };
        auto mma_a = [&](uint32_t dcol, uint32_t idesc, uint32_t boff, uint32_t brows, uint32_t acc) {     // A operand in TMEM
            umma_bf16_ts(dcol, acol, wd(boff), idesc, acc);
            umma_bf16_ts(dcol, acol + 8u, wd(boff + brows * 32), idesc, 1u);
        };
        auto mma_bias = [&](uint32_t dcol, uint32_t idesc, uint32_t boff, uint32_t row0, uint32_t acc) {
            umma_bf16(dcol, dOnes, wd(boff + row0 * 32), idesc, acc);
        };
        auto blk = [&](int j) { return kBlock0 + (uint32_t)j * kBlockBytes; };
        // condition tile of this slot's tile ts -> buffer ts & 1 (ycond viewed as [tokens * 256 pixel rows, 32 channels];
        // two 16-channel boxes = the two K chunks of the operand)
        auto load_y = [&](int ts) {
            int m, half;
            tile_of(k, ts, m, half);
            const int b = ts & 1;
            const uint32_t dst = sY + (uint32_t)(2 * k + b) * kYTile;
            const int row = m * 256 + half * 128;
            mbar_expect_tx(y_full(k, b), kYTile);
            tma_load_2d(dst, &ymap, y_full(k, b), 0, row);
            tma_load_2d(dst + 128 * 32, &ymap, y_full(k, b), 16, row);
        };
        if (elect_one()) {
            if (nt > 0) load_y(0);
            if (nt > 1) load_y(1);
        }
        __syncwarp();
        if (nt > 0) {           // prologue: scale' | gate of the first tile's block 0 (SC | GT are adjacent columns: N = 64)
            mbar_wait(y_full(k, 0), 0);
            tc_fence_after();
            if (elect_one()) {
                mma_bias(tcol + kColSC, id64, blk(0) + kBsg, 0, 0u);
                mma_y(tcol + kColSC, id64, dY0, blk(0) + kWsg, 64, 0, 1u);
                umma_commit(p_bar(k));
            }
            __syncwarp();
        }
        uint32_t aphase = 0;
        for (int ts = 0; ts < nt; ++ts) {
            const int yb = ts & 1;
            const uint64_t dy = dY0 + (uint64_t)(yb ? (kYTile >> 4) : 0);
            for (int s = 0; s < nstage; ++s) {
                mbar_wait(a_bar(k), aphase);
                aphase ^= 1;
                DTC_TRACE(1, 100 + s);
                const bool last = s == nstage - 1;
                const bool next = last && ts + 1 < nt;
                tc_fence_after();
                if (last) {
                    if (elect_one()) {
                        mma_a(tcol + kColH, id16, blk(R) + kWf, 16, 0u);
                        mma_bias(tcol + kColH, id16, blk(R) + kBf, 0, 1u);
                        umma_commit(d_bar(k));
                    }
                    __syncwarp();
                    if (next) {         // next tile's scale' | gate: SC and GT were last read two / zero stages ago
                        mbar_wait(y_full(k, yb ^ 1), (uint32_t)(((ts + 1) >> 1) & 1));
                        tc_fence_after();
                        if (elect_one()) {
                            const uint64_t dyn = dY0 + (uint64_t)(yb ? 0 : (kYTile >> 4));
                            mma_bias(tcol + kColSC, id64, blk(0) + kBsg, 0, 0u);
                            mma_y(tcol + kColSC, id64, dyn, blk(0) + kWsg, 64, 0, 1u);
                            umma_commit(p_bar(k));      // its own barrier: d_bar must never run two phases ahead of its waiters
                        }
                        __syncwarp();
                    }
                    if (ts + 2 < nt) {  // refill this tile's buffer (its last readers were committed two stages ago)
                        mbar_wait(y_empty(k, yb), (uint32_t)((ts >> 1) & 1));
                        if (elect_one()) load_y(ts + 2);
                        __syncwarp();
                    }
                } else if ((s & 1) == 0) {
                    const int j = s >> 1;
                    if (elect_one()) {
                        // H = 1/2 W0 h: the activation-dependent product last, so the accumulator's other addends are
                        // already in flight; the waiting epilogue needs only this group
                        mma_bias(tcol + kColH, id32, blk(j) + kB0, 0, 0u);
                        mma_y(tcol + kColH, id32, dy, blk(j) + kC0, 32, 0, 1u);
                        mma_a(tcol + kColH, id32, blk(j) + kW0, 32, 1u);
                        umma_commit(d_bar(k));
                        DTC_TRACE(1, 200 + s);
                        // not on the critical path (needed two stages later); covered by the next commit
                        if (j + 1 < R) {                    // SC_j has been read (this stage's arrival)
                            mma_bias(tcol + kColSC, id32, blk(j + 1) + kBsg, 0, 0u);
                            mma_y(tcol + kColSC, id32, dy, blk(j + 1) + kWsg, 64, 0, 1u);
                        }
                        if (j >= 1) {                       // GT_{j-1} has been read (same arrival)
                            mma_bias(tcol + kColGT, id32, blk(j) + kBsg, 32, 0u);
                            mma_y(tcol + kColGT, id32, dy, blk(j) + kWsg, 64, 32, 1u);
                        }
                        if (j == R - 1) umma_commit(y_empty(k, yb));        // last readers of this condition tile
                        DTC_TRACE(1, 300 + s);
                    }
                    __syncwarp();
                } else {
                    const int j = s >> 1;
                    if (elect_one()) {
                        mma_bias(tcol + kColH, id32, blk(j) + kB2, 0, 0u);
                        mma_a(tcol + kColH, id32, blk(j) + kW2, 32, 1u);
                        umma_commit(d_bar(k));
                        DTC_TRACE(1, 200 + s);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp < kEpiWarps) {
        // ================================================================== epilogue slots: thread = pixel = TMEM lane
        const int k = warp >> 2;
        const int l = (warp & 3) * 32 + lane;                   // lane inside the tile
        const uint32_t tcol = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)k * kSlotCols;
        const int nt = slot_tiles(k);
        const size_t plane = (size_t)P.H * P.W;
        const int L = P.Hp * P.Wp;
        const int Bimg = P.pair ? item_tokens / L : 0;          // images of the shared state (pair mode)
        uint32_t dphase = 0;
        uint32_t pphase = 0;
#ifdef DECO_DTC_TRACE
        int trace_n = 0;
        const bool trace_on = k == 0 && l == 0;
#endif
        auto wait_d = [&]() { mbar_wait(d_bar(k), dphase); dphase ^= 1; tc_fence_after(); };
        auto release = [&]() {                                  // TMEM stores / loads of this warp are complete and fenced
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_bar(k));
        };
        float g_ = P.g, dt_ = P.dt, c0_ = P.c0, c1_ = P.c1;
        if (P.pair && P.dev) { g_ = __ldg(P.dev); dt_ = __ldg(P.dev + 1); c0_ = __ldg(P.dev + 2); c1_ = __ldg(P.dev + 3); }
        float u_keep[3] = {0.f, 0.f, 0.f};
        // pixel of this thread in tile ts: offsets of its image row (x / out) and inside the plane
        auto locate = [&](int ts, size_t& xoff, size_t& ooff) {
            int m, half;
            tile_of(k, ts, m, half);
            const int pix = half * 128 + l;                     // pixel inside the patch: ky = pix / 16, kx = pix % 16
            const int img_row = m / L;                          // row of the CFG batch
            const int tok = m - img_row * L;
            const int py = tok / P.Wp, px = tok - py * P.Wp;
            const size_t pix_off = (size_t)(py * 16 + (pix >> 4)) * P.W + (size_t)px * 16 + (pix & 15);
            const int xrow = P.pair ? (img_row >= Bimg ? img_row - Bimg : img_row) : img_row;
            xoff = (size_t)xrow * 3 * plane + pix_off;
            ooff = (size_t)img_row * 3 * plane + pix_off;
        };
        // the image samples of a tile are fetched one tile ahead (plain loads: pair mode rewrites x in place later)
        float rgb[3] = {0.f, 0.f, 0.f}, rgb_next[3] = {0.f, 0.f, 0.f};
        size_t xoff = 0, ooff = 0, xoff_next = 0, ooff_next = 0;
        if (nt > 0) {
            locate(0, xoff_next, ooff_next);
#pragma unroll
            for (int c = 0; c < 3; ++c) rgb_next[c] = P.x[xoff_next + c * plane];
        }
        for (int ts = 0; ts < nt; ++ts) {
            DTC_TRACE(0, 1);
            const int pix = (((first + ((P.pair ? (ts >> 1) : ts) * kSlots + k) * step) & 1) << 7) + l;
            xoff = xoff_next; ooff = ooff_next;
            if (!(P.pair && (ts & 1))) {                        // pair mode: the cond tile re-uses the uncond tile's samples
#pragma unroll
                for (int c = 0; c < 3; ++c) rgb[c] = rgb_next[c];
            }
            // ---- x = T'[pixel] + W' bf16(rgb)   (NerfEmbedder + input_proj, fp32)
            float x[kHx];
            {
                const float r0 = round_bf(rgb[0]), r1 = round_bf(rgb[1]), r2 = round_bf(rgb[2]);
                const float4* trow = reinterpret_cast<const float4*>(sTab + pix * kTabPitch);
                const float4* wr = reinterpret_cast<const float4*>(sWrgb);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 tv = trow[q];
                    const float tt[4] = {tv.x, tv.y, tv.z, tv.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float4 w = wr[4 * q + e];
                        x[4 * q + e] = fmaf(w.x, r0, fmaf(w.y, r1, fmaf(w.z, r2, tt[e])));
                    }
                }
            }
            // LayerNorm of x times `mul` (per channel, 32 raw fp32 words or null = 1) -> bf16 A operand
            auto norm_to_a = [&](const uint32_t* mul) {
                // single pass: sum and sum of squares (|mean| is a few sigma at most here: the cancellation in
                // E[x^2] - mean^2 stays ~1e-6 relative, far below the bf16 rounding of the result)
                float sm = 0.f, sq = 0.f;
#pragma unroll
                for (int c = 0; c < kHx; ++c) { sm += x[c]; sq = fmaf(x[c], x[c], sq); }
                const float mean = sm * (1.0f / kHx);
                const float var = fmaxf(fmaf(-mean, mean, sq * (1.0f / kHx)), 0.f);
                const float r = rsqrtf(var + 1e-6f);
                const float nmr = -mean * r;
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float a = fmaf(x[2 * i], r, nmr), b = fmaf(x[2 * i + 1], r, nmr);
                    if (mul) { a *= __uint_as_float(mul[2 * i]); b *= __uint_as_float(mul[2 * i + 1]); }
                    pk[i] = pack_bf2(a, b);
                }
                tmem_st16(tcol + kColA, pk);
                tmem_st_wait();
            };
            // ---- stage 0: h~ = LN(x) . sc'   (scale' | gate of block 0 were issued with the previous tile's final layer)
            DTC_TRACE(0, 2);
            mbar_wait(p_bar(k), pphase);
            pphase ^= 1;
            tc_fence_after();
            DTC_TRACE(0, 3);
            {
                uint32_t sc[32];
                tmem_ld32(tcol + kColSC, sc);
                tmem_ld_wait();
                DTC_TRACE(0, 4);
                norm_to_a(sc);
                DTC_TRACE(0, 5);
                release();
                DTC_TRACE(0, 6);
            }
            if (ts + 1 < nt) {      // next tile's pixel: its global loads fly while this tile runs
                locate(ts + 1, xoff_next, ooff_next);
                if (!(P.pair && !(ts & 1))) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) rgb_next[c] = P.x[xoff_next + c * plane];
                }
            }
            for (int j = 0; j < R; ++j) {
                // ---- stage 2j + 1: u = silu(H) with H = v / 2 -> A operand
                wait_d();
                DTC_TRACE(0, 10);
                {
                    uint32_t h[32];
                    tmem_ld32(tcol + kColH, h);
                    tmem_ld_wait();
                    DTC_TRACE(0, 11);
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float a = __uint_as_float(h[2 * i]), b = __uint_as_float(h[2 * i + 1]);
                        pk[i] = pack_bf2(fmaf(a, tanh_approx(a), a), fmaf(b, tanh_approx(b), b));
                    }
                    tmem_st16(tcol + kColA, pk);
                    tmem_st_wait();
                    DTC_TRACE(0, 12);
                    release();
                    DTC_TRACE(0, 13);
                }
                // ---- stage 2j + 2: x += gate . H2, then the next block's modulated norm (or the final norm)
                wait_d();
                DTC_TRACE(0, 20);
                {
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        uint32_t gt[16], h2[16];
                        tmem_ld16(tcol + kColGT + (uint32_t)(hf * 16), gt);
                        tmem_ld16(tcol + kColH + (uint32_t)(hf * 16), h2);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            x[hf * 16 + i] = fmaf(__uint_as_float(gt[i]), __uint_as_float(h2[i]), x[hf * 16 + i]);
                    }
                    if (j + 1 < R) {
                        uint32_t sc[32];
                        tmem_ld32(tcol + kColSC, sc);
                        tmem_ld_wait();
                        norm_to_a(sc);
                    } else {
                        norm_to_a(nullptr);
                    }
                    DTC_TRACE(0, 21);
                    release();
                    DTC_TRACE(0, 22);
                }
            }
            // ---- output stage: 3 channels of the final linear
            wait_d();
            DTC_TRACE(0, 30);
            float o[3];
            {
                uint32_t f[8];
                tmem_ld8(tcol + kColH, f);
                tmem_ld_wait();
                o[0] = __uint_as_float(f[0]); o[1] = __uint_as_float(f[1]); o[2] = __uint_as_float(f[2]);
            }
            if (!P.pair) {
                if (P.out_bf16) {
                    __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(P.out) + ooff;
#pragma unroll
                    for (int c = 0; c < 3; ++c) op[c * plane] = f2bf(o[c]);
                } else {
                    float* op = reinterpret_cast<float*>(P.out) + ooff;
#pragma unroll
                    for (int c = 0; c < 3; ++c) op[c * plane] = o[c];
                }
            } else if ((ts & 1) == 0) {
#pragma unroll
                for (int c = 0; c < 3; ++c) u_keep[c] = o[c];
            } else {
                // guidance + multistep update of the fp32 state (csrc/sampler.cu, same expression order)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const size_t ob = xoff + c * plane;
                    const float pred = fmaf(g_, o[c] - u_keep[c], u_keep[c]);
                    float v = c0_ * pred;
                    if (P.p1) v = fmaf(c1_, P.p1[ob], v);
                    const float xn = fmaf(dt_, v, P.x_base ? P.x_base[ob] : rgb[c]);
                    P.x_out[ob] = xn;
                    if (P.pred_out) P.pred_out[ob] = pred;
                    if (P.u8_out) P.u8_out[ob] = to_u8(xn);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kEpiWarps) tmem_dealloc(tmem, 512);
}

}  // namespace dtc
}  // namespace deco

using namespace deco;

extern "C" int deco_decoder_tc_blob_bytes(int num_res_blocks) { return (int)dtc::blob_bytes(num_res_blocks); }

// ysilu: bf16 [tokens, 256 * 32] = silu(cond_embed(s)) (the cond_embed GEMM with the SiLU epilogue); blob: packed by
// deco_b200/denoiser.py::pack_decoder_tc.  pair = 0: out [rows, 3, H, W] = decoder(x rows).  pair != 0: rows = 2B stacked
// [uncond || cond] over the SAME image state x [B, 3, H, W]; the guided multistep update is applied to x_base (NULL = x)
// and written to x_out.
extern "C" int deco_pixel_decoder_tc(const float* x, const void* ysilu_bf16, const void* blob, void* out, int out_is_bf16,
                                     int rows, int H, int W, int patch, int hidden_x, int num_res_blocks,
                                     int pair, const float* dev_scalars, float g, float dt, float c0, float c1,
                                     const float* x_base, const float* p1, float* x_out, float* pred_out, void* u8_out,
                                     void* stream)
{
    using namespace deco::dtc;
    DECO_CHECK_ARG(x && ysilu_bf16 && blob, "pixel_decoder_tc: null pointer");
    if (patch != 16 || hidden_x != kHx) {
        deco_set_error("pixel_decoder_tc: built for patch_size 16 and hidden_size_x 32 (got %d, %d)", patch, hidden_x);
        return DECO_ERR_UNSUPPORTED;
    }
    DECO_CHECK_ARG(rows > 0 && H % 16 == 0 && W % 16 == 0 && num_res_blocks >= 1 && num_res_blocks <= kMaxR,
                   "pixel_decoder_tc: bad shape rows=%d H=%d W=%d R=%d", rows, H, W, num_res_blocks);
    DECO_CHECK_ARG(pair ? (rows % 2 == 0 && x_out != nullptr) : (out != nullptr), "pixel_decoder_tc: missing output");
    DECO_CHECK_ARG((((uintptr_t)ysilu_bf16 | (uintptr_t)blob) & 15) == 0, "pixel_decoder_tc: pointers must be 16-byte aligned");
    Params P = {};
    P.x = x; P.blob = blob; P.out = out; P.out_bf16 = out_is_bf16;
    P.R = num_res_blocks; P.H = H; P.W = W; P.Hp = H / 16; P.Wp = W / 16;
    const long long L = (long long)P.Hp * P.Wp;
    P.tokens = (long long)rows * L;
    P.pair = pair ? 1 : 0;
    P.tokens_half = P.tokens / 2;
    P.dev = dev_scalars; P.g = g; P.dt = dt; P.c0 = c0; P.c1 = c1;
    P.x_base = x_base; P.p1 = p1; P.x_out = x_out; P.pred_out = pred_out; P.u8_out = (uint8_t*)u8_out;
    DECO_CHECK_ARG(P.tokens * 256 < (1LL << 31), "pixel_decoder_tc: too many pixel rows for one launch (%lld tokens)", P.tokens);

    PFN_encodeTiled enc = get_tensormap_encoder();
    if (!enc) { deco_set_error("cuTensorMapEncodeTiled entry point not available"); return DECO_ERR_DRIVER; }
    CUtensorMap ymap;
    cuuint64_t dims[2] = {(cuuint64_t)kHx, (cuuint64_t)(P.tokens * 256)};
    cuuint64_t strides[1] = {(cuuint64_t)kHx * 2};
    cuuint32_t box[2] = {16, 128};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&ymap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ysilu_bf16), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { deco_set_error("pixel_decoder_tc: cuTensorMapEncodeTiled failed: %d", (int)r); return DECO_ERR_DRIVER; }

    const int smem = (int)smem_bytes(num_res_blocks);
    static unsigned long long attr_done = 0;
    static int attr_smem = 0;
    if (!device_setup_done(attr_done) || smem > attr_smem) {
        cudaError_t e = cudaFuncSetAttribute(pixel_decoder_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) { deco_set_error("pixel_decoder_tc attr: %s", cudaGetErrorString(e)); return (int)e; }
        mark_device_setup(attr_done);
        attr_smem = 227 * 1024;
    }
    const long long items = (pair ? P.tokens_half : P.tokens) * 2;
    long long grid = device_sm_count();
    if (grid > items) grid = items;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = deco_pdl_enabled() ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, pixel_decoder_tc_kernel, ymap, P);
    if (e != cudaSuccess) { deco_set_error("pixel_decoder_tc launch failed: %s", cudaGetErrorString(e)); return (int)e; }
    return DECO_OK;
}

#ifdef DECO_DTC_TRACE
extern "C" int deco_dtc_trace_copy(long long* host_buf) {
    return (int)cudaMemcpyFromSymbol(host_buf, deco::dtc::g_dtc_trace, sizeof(long long) * 2 * 4096);
}
#endif
