// tcgen05 / TMEM / TMA GEMM for the prefill-sized linears:  C[M,N] = A[M,K] . B[N,K]^T (+bias)
//   kind BF16 : A, B bf16 (K-major, as the reference stores activations [T,in] and weights [out,in])
//               replaces xllm::kernel::cuda::matmul -> F::linear (kernels/cuda/matmul.cpp:20-24)
//   kind FP8  : A, B e4m3, C = a_scale * (b_scale * acc) + bias, per-tensor or per-token / per-channel scales
//               replaces cutlass_scaled_mm (cutlass_w8a8/scaled_mm_entry.cu:55-108, c3x/scaled_mm_sm100_fp8_dispatch.cuh)
//   kind W4   : B is the tile-packed int4 weight of linear_small_m.cu; 4 converter warps dequantise
//               bf16((q-z)*s) (bit-exact with the spec) straight into the 128B-swizzled K-major stage buffer the
//               MMA reads, so the unpack is fused into the tcgen05 main loop (no bf16 copy of W ever exists in HBM).
//
// One persistent CTA per SM, warp-specialised:
//   warp 0      TMA producer (A tiles, and B tiles / packed-B bulk copies), one elected lane
//   warp 1      MMA issuer: tcgen05.mma.cta_group::1, 128 x BLOCK_N x 16 (x32 for fp8) per instruction, one elected lane;
//               owns the TMEM allocation (2 accumulator stages so the epilogue of tile i overlaps tile i+1)
//   warps 2..5  epilogue: tcgen05.ld 32x32b -> scale/bias -> bf16 -> global   (warp w owns TMEM lanes 32*(w%4)..)
//   warps 6..9  (W4 only) converters: packed smem -> LOP3/HSUB2/HMUL2 -> swizzled bf16 smem -> fence.proxy.async
// Pipelines are mbarrier rings: smem full/empty (TMA <-> MMA), packed full / B ready (TMA <-> converters <-> MMA),
// TMEM full/empty (MMA <-> epilogue).
#include "tc_common.cuh"

namespace xb {
namespace tc {

EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_bytes, uint32_t box_rows,
                 uint32_t box_cols, int elem_bytes) {
  EncodeTiledFn enc = get_encode_tiled();
  XB_CHECK(enc != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8;
  CUresult r = enc(out, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  XB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (rows %llu cols %llu pitch %llu)", (int)r,
           (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)pitch_bytes);
  return 0;
}

int make_tmap_2d_raw(CUtensorMap* out, const void* base, int dtype, uint64_t rows, uint64_t cols, uint64_t pitch_bytes,
                     uint32_t box_rows, uint32_t box_cols, int swizzle) {
  EncodeTiledFn enc = get_encode_tiled();
  XB_CHECK(enc != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, (CUtensorMapDataType)dtype, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, (CUtensorMapSwizzle)swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  XB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(raw) failed with CUresult %d", (int)r);
  return 0;
}

int make_tmap_3d_bf16(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                      uint64_t stride2_bytes, uint32_t b0, uint32_t b1, uint32_t b2) {
  EncodeTiledFn enc = get_encode_tiled();
  XB_CHECK(enc != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {b0, b1, b2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  XB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(3d) failed with CUresult %d", (int)r);
  return 0;
}

enum { kKindBF16 = 0, kKindFP8 = 1, kKindW4 = 2, kKindW8 = 3 };
constexpr bool kind_is_wq(int kind) { return kind == kKindW4 || kind == kKindW8; }   // weight-only: packed B + converter warps

struct GemmParams {
  __nv_bfloat16* c;
  int64_t ldc;
  const __nv_bfloat16* bias;   // [N] or null
  const float* a_scale;        // fp8: [1] or [M]
  const float* b_scale;        // fp8: [1] or [N]
  int a_scale_per_row, b_scale_per_col;
  // W4
  const uint4* qweight;        // tile-packed [N/16][K/64][32]
  const uint32_t* meta;        // [K/g][N]
  int gshift;                  // log2(k64 tiles per quantisation group)
  int M, N, K;
  int use_tma_store;           // C rows are 16-byte aligned: epilogue stages through smem and TMA-stores
  // swap-AB (decode-sized M): the WEIGHT rows ride in the 128-row MMA M slot and the tokens are the N tile, so every CTA
  // streams 128 weight rows instead of idling 3/4 of a 128-token tile.  M / N / a_scale / b_scale are already in the
  // swapped roles (M = output channels, N = tokens, a_scale per channel, b_scale per token); the epilogue indexes bias by
  // row, keeps the reference's product order token_scale * (channel_scale * acc) and stores C transposed: c[col][row]
  int swap_ab;
  // split-K (swap-AB FP8 only): a decode-sized GEMM has N/128 tiles - 10 for a TP8 qkv shard - so most SMs would idle while a
  // few CTAs stream all of K.  Each output tile is cut into split_k k-ranges walked by different CTAs; every CTA leaves its
  // fp32 partial in splitk_ws, takes a ticket (self-resetting atomicInc, so graph replays need no memset), and the LAST
  // arriver sums the partials in split order (deterministic) and runs the epilogue.  Nobody waits, so no co-residency
  // requirement.  Workspace: caller-owned (xb_set_gemm_splitk_workspace).
  int split_k;
  float* splitk_ws;            // [tiles][split_k][kBlockN][128] fp32
  unsigned int* splitk_tickets;  // [tiles], zero at rest
  int debug_skip_convert;      // XB_GEMM_DEBUG_SKIP_CONVERT=1 (timing diagnosis only, results are garbage): converters only hand the stage over
};

constexpr int kBlockM = 128;
constexpr int kNumEpiWarps = 4;

template <int kKind, int kBlockN, int kMT = 1 /* 128-row M tiles per CTA sharing one B stage */,
          int kCG = 1 /* 2: CTA pair (cta_group::2): 256 x kBlockN tile, each CTA holds / converts kBlockN/2 rows of B */>
struct GemmCfg {
  static_assert(kCG == 1 || (kCG == 2 && kMT == 1), "the CTA-pair kernel takes one 128-row M tile per CTA");
  static constexpr int kCtaM = kBlockM * kMT;
  static constexpr int kCtaN = kBlockN / kCG;                       // B rows staged (and dequantised) by ONE CTA
  static constexpr int kElemA = kKind == kKindFP8 ? 1 : 2;
  static constexpr int kBlockK = 128 / kElemA;                     // one 128-byte swizzle row per tile row
  static constexpr int kUmmaK = 32 / kElemA;                       // 16 (bf16) / 32 (fp8) elements = 32 bytes
  static constexpr int kABytes = kCtaM * 128;
  static constexpr int kBBytes = kCtaN * 128;
  static constexpr int kPackedBytes = kKind == kKindW4 ? kCtaN * kBlockK / 2 : (kKind == kKindW8 ? kCtaN * kBlockK : 0);   // int4 / int8 tile
  static constexpr int kMetaBytes = kind_is_wq(kKind) ? kCtaN * 4 : 0;
  // BF16 / FP8: a stage holds the A and B tiles.  W4: a stage holds A + the PACKED int4 B tile + its scale/zero words
  // (small, so the TMA ring can be deep enough to cover HBM latency) and the dequantised bf16 B lives in a separate
  // 2-deep ring written by the converter warps.
  static constexpr int kPackedRegion = ((kPackedBytes + 1023) / 1024) * 1024;      // stages stay 1024-byte aligned (SW128 A tiles)
  static constexpr int kStageBytes = kind_is_wq(kKind) ? kABytes + kPackedRegion + ((kMetaBytes + 1023) / 1024) * 1024
                                                      : kABytes + kBBytes;
  static_assert(kStageBytes % 1024 == 0 && kBBytes % 1024 == 0, "128B-swizzled tiles need 1024-byte aligned stages");
  // dequantised-B ring depth.  CTA pairs: the converters <-> MMA hand-off crosses SMs (multicast commit one way, remote
  // mbarrier arrive the other: a few hundred cycles each), so the converters must run further ahead of the MMA
  static constexpr int kBStages = kind_is_wq(kKind) ? (kCG == 2 ? 4 : 2) : 0;
  static constexpr int kBRing = kBStages > 0 ? kBStages : 2;        // barriers of the ring (two dummies when unused)
  // converters: 8 warps convert one B stage (two per SM sub-partition).  A converter's k block is a latency chain (shared
  // loads -> ALU -> shared stores -> proxy fence -> arrive, ~2/3 of the 512 MMA cycles of a pair k block: ncu showed the
  // converter warps parked at the fence and the MMA issuer asleep on B-ready, tensor pipe 61 %), so CTA pairs run TWO sets
  // of 8 warps that take alternate k blocks: each set has two MMA periods per stage
  static constexpr int kConvSets = kind_is_wq(kKind) ? (kCG == 2 ? 2 : 1) : 0;
  static constexpr int kConvPerStage = kind_is_wq(kKind) ? 8 : 0;
  static constexpr int kConvWarps = kConvSets * kConvPerStage;
  static constexpr int kEpiStageBytes = kNumEpiWarps * 2 * 4096;   // per epilogue warp: 2 x [32 rows x 64 cols] bf16
  static constexpr int kBudget = 227 * 1024 - 1024 - 512 - kEpiStageBytes - kBStages * kBBytes;
  static constexpr int kStages = kBudget / kStageBytes > 8 ? 8 : kBudget / kStageBytes;
  static constexpr int kThreads = (2 + kNumEpiWarps + kConvWarps) * 32;
  static constexpr int kAccCols = kMT * kBlockN;                     // one accumulator stage
  static constexpr int kTmemUsed = 2 * kAccCols;                   // two accumulator stages
  static_assert(kTmemUsed <= 512, "TMEM has 512 columns");
  static constexpr int kTmemCols = kTmemUsed <= 32 ? 32 : kTmemUsed <= 64 ? 64 : kTmemUsed <= 128 ? 128 : kTmemUsed <= 256 ? 256 : 512;   // allocations are powers of two
  static constexpr int kSmemBytes = kStages * kStageBytes + kBStages * kBBytes + kEpiStageBytes + 1024 /*align*/ + 512 /*barriers*/;
};

template <int kKind, int kBlockN, int kMT, int kCG>
__global__ void __launch_bounds__(GemmCfg<kKind, kBlockN, kMT, kCG>::kThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_m,
                    const GemmParams p) {
  using Cfg = GemmCfg<kKind, kBlockN, kMT, kCG>;
  constexpr int kStages = Cfg::kStages;
  constexpr bool kPair = kCG == 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* bbuf = smem + kStages * Cfg::kStageBytes;                  // W4: [kBStages][kBBytes] dequantised B ring
  uint8_t* epi_stage = bbuf + Cfg::kBStages * Cfg::kBBytes;           // [kNumEpiWarps][2][4096]
  uint8_t* bar_mem = epi_stage + Cfg::kEpiStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_mem);          // [kStages] TMA -> MMA (A, and B for BF16/FP8)
  uint64_t* empty_bar = full_bar + kStages;                           // [kStages] MMA -> TMA
  uint64_t* packed_bar = empty_bar + kStages;                         // [kStages] TMA -> converters (W4)
  uint64_t* bready_bar = packed_bar + kStages;                        // [kBRing] converters -> MMA (W4 B ring)
  uint64_t* bempty_bar = bready_bar + Cfg::kBRing;                    // [kBRing] MMA -> converters
  uint64_t* tmem_full = bempty_bar + Cfg::kBRing;                     // [2]
  uint64_t* tmem_empty = tmem_full + 2;                               // [2]
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // CTA pair: rank 0 (leader) issues the MMAs and owns the barriers both CTAs feed (full, B-ready, TMEM-empty); a work
  // unit is one 256 x kBlockN tile, walked by both CTAs of the pair in the same order
  const uint32_t rank = kPair ? cluster_ctarank() : 0u;
  const int unit0 = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int unit_stride = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int m_blocks = (p.M + Cfg::kCtaM * kCG - 1) / (Cfg::kCtaM * kCG);
  const int n_blocks = (p.N + kBlockN - 1) / kBlockN;
  const int num_kb = (p.K + Cfg::kBlockK - 1) / Cfg::kBlockK;
  // split-K exists in the swap-AB FP8 instantiations only (compile-time off elsewhere: no extra registers in the big tiles)
  constexpr bool kSplitOK = kKind == kKindFP8 && kBlockN <= 64 && kCG == 1 && kMT == 1;
  const int split_k = kSplitOK ? p.split_k : 1;
  const int kb_per = (num_kb + split_k - 1) / split_k;
  const int num_tiles = m_blocks * n_blocks * split_k;      // work units: (output tile, k range); the splits of a tile are adjacent

  auto stage_a = [&](int s) { return smem + s * Cfg::kStageBytes; };
  auto stage_b = [&](int s) { return smem + s * Cfg::kStageBytes + Cfg::kABytes; };               // BF16 / FP8
  auto stage_packed = [&](int s) { return smem + s * Cfg::kStageBytes + Cfg::kABytes; };          // W4
  auto stage_meta = [&](int s) { return smem + s * Cfg::kStageBytes + Cfg::kABytes + Cfg::kPackedRegion; };
  auto b_ring = [&](int bs) { return bbuf + bs * Cfg::kBBytes; };

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar + s, 1);
      mbar_init(empty_bar + s, 1);
      mbar_init(packed_bar + s, 1);
    }
    for (int s = 0; s < Cfg::kBRing; ++s) {
      mbar_init(bready_bar + s, Cfg::kConvPerStage > 0 ? Cfg::kConvPerStage * kCG : 1);   // one elected arrive per converter warp of the stage's set (of both CTAs)
      mbar_init(bempty_bar + s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tmem_full + s, 1);
      mbar_init(tmem_empty + s, kNumEpiWarps * kCG);
    }
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (kind_is_wq(kKind)) tma_prefetch_desc(&tmap_m);
    if (p.use_tma_store) tma_prefetch_desc(&tmap_c);
  }
  if (warp == 1) {
    if (kPair) tmem_alloc_cg2(tmem_base_smem, Cfg::kTmemCols);
    else tmem_alloc(tmem_base_smem, Cfg::kTmemCols);
  }
  tc_fence_before_sync();
  if (kPair) cluster_sync_all();   // the peer's barriers are initialised before any remote arrive / complete_tx / commit
  else __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_base_smem;
  // the leader's copies of the shared barriers (shared::cluster addresses; for rank 0 these are its own)
  const uint32_t full_bar_leader = kPair ? cluster_map(smem_u32(full_bar), 0) : smem_u32(full_bar);
  const uint32_t bready_leader = kPair ? cluster_map(smem_u32(bready_bar), 0) : smem_u32(bready_bar);
  const uint32_t tmem_empty_leader = kPair ? cluster_map(smem_u32(tmem_empty), 0) : smem_u32(tmem_empty);

  pdl_launch_dependents();
  pdl_wait();   // A (activations) comes from the producer kernel

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = unit0; tile < num_tiles; tile += unit_stride) {
        const int otile = tile / split_k, ks = tile - otile * split_k;
        const int m_blk = otile % m_blocks, n_blk = otile / m_blocks;
        const int a_row = (m_blk * kCG + (int)rank) * Cfg::kCtaM;           // this CTA's 128 (x kMT) rows of A
        const int b_row = n_blk * kBlockN + (int)rank * Cfg::kCtaN;         // this CTA's share of the B rows
        const int kb0 = ks * kb_per, kb1 = min(num_kb, kb0 + kb_per);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar + s, ph ^ 1);
          if (kind_is_wq(kKind)) {
            // packed B: the kCtaN/16 row tiles of this k tile are one 2-D TMA box (tensor [N/16][K/64 * 128 words],
            // box [kCtaN/16][128 words], no swizzle) - one instruction instead of kCtaN/16 bulk copies - plus the
            // group's scale/zero words for these kCtaN rows (tensor [K/g][N], box [1][kCtaN]); both stay CTA-local
            mbar_expect_tx(packed_bar + s, Cfg::kPackedBytes + Cfg::kMetaBytes);
            // (W8: one byte per weight - 256 words per row tile and k tile instead of 128)
            tma_load_2d(stage_packed(s), &tmap_b, packed_bar + s, kb * (kKind == kKindW8 ? 256 : 128), b_row / 16);
            tma_load_2d(stage_meta(s), &tmap_m, packed_bar + s, b_row, kb >> p.gshift);
            if (kPair) {
              if (rank == 0) mbar_expect_tx(full_bar + s, Cfg::kABytes * 2);   // both CTAs' A tiles complete on the leader
              tma_load_2d_cg2(stage_a(s), &tmap_a, full_bar_leader + s * 8, kb * Cfg::kBlockK, a_row);
            } else {
              mbar_expect_tx(full_bar + s, Cfg::kABytes);
              tma_load_2d(stage_a(s), &tmap_a, full_bar + s, kb * Cfg::kBlockK, a_row);
            }
          } else if (kPair) {
            if (rank == 0) mbar_expect_tx(full_bar + s, (Cfg::kABytes + Cfg::kBBytes) * 2);
            tma_load_2d_cg2(stage_a(s), &tmap_a, full_bar_leader + s * 8, kb * Cfg::kBlockK, a_row);
            tma_load_2d_cg2(stage_b(s), &tmap_b, full_bar_leader + s * 8, kb * Cfg::kBlockK, b_row);
          } else {
            mbar_expect_tx(full_bar + s, Cfg::kABytes + Cfg::kBBytes);
            tma_load_2d(stage_a(s), &tmap_a, full_bar + s, kb * Cfg::kBlockK, a_row);
            tma_load_2d(stage_b(s), &tmap_b, full_bar + s, kb * Cfg::kBlockK, b_row);
          }
          if (++s == kStages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ======================= MMA issuer (leader CTA of a pair) =======================
    constexpr uint32_t idesc = kKind == kKindFP8 ? umma_idesc(0, 0, kBlockM * kCG, kBlockN) : umma_idesc(1, 1, kBlockM * kCG, kBlockN);
    int s = 0;
    uint32_t ph = 0;
    int as = 0;
    uint32_t aph = 0;
    int bs = 0;
    uint32_t bph = 0;
    for (int tile = unit0; tile < num_tiles; tile += unit_stride) {
      // (pairs: BOTH CTAs' epilogues have drained this accumulator stage.  The wait is the plain CTA-scope one even for
      // barriers the peer arrives on remotely: the waiting thread reads nothing the peer wrote through the generic proxy -
      // it only issues MMAs - and a cluster-scope acquire costs an L1 invalidation (CCTL.IVALL) per k block, which
      // halved the W4 pair kernel: 600 vs 1020 TF/s, profiles/r02c_gemm_w4_pair.md)
      mbar_wait(tmem_empty + as, aph ^ 1);     // epilogue has drained this accumulator stage
      tc_fence_after_sync();
      const uint32_t d_tmem = tmem_base + as * Cfg::kAccCols;
      const int ks = tile % split_k;
      const int kb0 = ks * kb_per, kb1 = min(num_kb, kb0 + kb_per);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(full_bar + s, ph);
        if (kind_is_wq(kKind)) mbar_wait(bready_bar + bs, bph);
        tc_fence_after_sync();
        if (lane == 0) {
          const uint32_t a_addr = smem_u32(stage_a(s));
          const uint32_t b_addr = smem_u32(kind_is_wq(kKind) ? b_ring(bs) : stage_b(s));
#pragma unroll
          for (int k = 0; k < Cfg::kBlockK / Cfg::kUmmaK; ++k) {
            const uint64_t db = umma_desc_sw128(b_addr + k * 32);
#pragma unroll
            for (int mt = 0; mt < kMT; ++mt) {       // the M tiles of this CTA reuse the same B stage
              const uint64_t da = umma_desc_sw128(a_addr + mt * (kBlockM * 128) + k * 32);
              if (kPair) {
                // M = 256 over the pair: the same shared-memory offsets address this CTA's and the peer's A rows / B rows
                if (kKind == kKindFP8) umma_f8_cg2(d_tmem, da, db, idesc, ((kb - kb0) | k) != 0);
                else umma_f16_cg2(d_tmem, da, db, idesc, ((kb - kb0) | k) != 0);
              } else if (kKind == kKindFP8) umma_f8(d_tmem + mt * kBlockN, da, db, idesc, ((kb - kb0) | k) != 0);
              else umma_f16(d_tmem + mt * kBlockN, da, db, idesc, ((kb - kb0) | k) != 0);
            }
          }
          if (kPair) {
            umma_commit_cg2(empty_bar + s);            // multicast: both CTAs' producers may refill the slot
            if (kind_is_wq(kKind)) umma_commit_cg2(bempty_bar + bs);
            if (kb == kb1 - 1) umma_commit_cg2(tmem_full + as);
          } else {
            umma_commit(empty_bar + s);                 // smem slot is free once these MMAs have read it
            if (kind_is_wq(kKind)) umma_commit(bempty_bar + bs);
            if (kb == kb1 - 1) umma_commit(tmem_full + as);
          }
        }
        __syncwarp();
        if (++s == kStages) { s = 0; ph ^= 1; }
        if (kind_is_wq(kKind) && ++bs == Cfg::kBRing) { bs = 0; bph ^= 1; }
      }
      if (++as == 2) { as = 0; aph ^= 1; }
    }
  } else if (warp >= 2 && warp < 2 + kNumEpiWarps) {
    // ======================= epilogue =======================
    // TMEM -> registers (tcgen05.ld 32x32b: thread = accumulator row) -> scale/bias -> bf16 -> 128B-swizzled staging
    // tile in smem -> TMA store (clips ragged M / N edges).  Falls back to direct global stores when C rows are not
    // 16-byte aligned.
    const int q = warp & 3;                     // TMEM lane quarter this warp may read
    const int ew = warp - 2;
    uint8_t* my_stage = epi_stage + ew * 2 * 4096;
    int sbuf = 0;
    int as = 0;
    uint32_t aph = 0;
    for (int tile = unit0; tile < num_tiles; tile += unit_stride) {
      const int otile = tile / split_k;
      const int m_blk = otile % m_blocks, n_blk = otile / m_blocks;
      mbar_wait(tmem_full + as, aph);
      tc_fence_after_sync();
      bool reduce_here = false;          // split-K: this CTA arrived last and owns the tile's epilogue
      const float* ws_tile = nullptr;
      if (kSplitOK && split_k > 1) {
        // leave this k range's fp32 partial in the workspace: [split][column][row] so a warp writes 128 contiguous bytes
        const int ks = tile - otile * split_k;
        float* wp = p.splitk_ws + ((size_t)otile * split_k + ks) * (kBlockN * kBlockM) + q * 32 + lane;
#pragma unroll 1
        for (int c = 0; c < kBlockN / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + as * Cfg::kAccCols + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) __stcg(wp + (c * 32 + j) * kBlockM, __uint_as_float(r[j]));
        }
        __threadfence();
        volatile uint32_t* flag = reinterpret_cast<volatile uint32_t*>(epi_stage);
        asm volatile("bar.sync 1, %0;" ::"n"(kNumEpiWarps * 32) : "memory");
        if (ew == 0 && lane == 0) *flag = atomicInc(p.splitk_tickets + otile, (unsigned)split_k - 1u);
        asm volatile("bar.sync 1, %0;" ::"n"(kNumEpiWarps * 32) : "memory");
        reduce_here = *flag == (uint32_t)split_k - 1u;
        if (reduce_here) {
          __threadfence();
          ws_tile = p.splitk_ws + (size_t)otile * split_k * (kBlockN * kBlockM) + q * 32 + lane;
        }
      }
      if (!(kSplitOK && split_k > 1) || reduce_here) {
#pragma unroll 1
      for (int mt = 0; mt < kMT; ++mt) {
      const int row0 = (m_blk * kCG + (int)rank) * Cfg::kCtaM + mt * kBlockM + q * 32;
      const int row = row0 + lane;
      float a_s = 1.f, b_s_uniform = 1.f;
      if (kKind == kKindFP8) {
        a_s = p.a_scale[p.a_scale_per_row ? min(row, p.M - 1) : 0];
        b_s_uniform = p.b_scale[0];
      }
#pragma unroll 1
      for (int c = 0; c < kBlockN / 32; ++c) {
        uint32_t r[32];
        if (kSplitOK && split_k > 1) {
          // sum the partials in split order (the same order whichever CTA arrived last: bit-reproducible)
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__ldcg(ws_tile + (c * 32 + j) * kBlockM));
#pragma unroll 1
          for (int sp = 1; sp < split_k; ++sp) {
            const float* wsp = ws_tile + (size_t)sp * (kBlockN * kBlockM);
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __ldcg(wsp + (c * 32 + j) * kBlockM));
          }
        } else {
          tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + as * Cfg::kAccCols + mt * kBlockN + c * 32, r);
          tmem_ld_wait();
        }
        const int col0 = n_blk * kBlockN + c * 32;
        uint32_t o[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          float x0 = __uint_as_float(r[j]), x1 = __uint_as_float(r[j + 1]);
          const int col = col0 + j;
          if (kKind == kKindFP8) {
            // per-tensor scale: one load per tile row (b_s_uniform), not two global loads per element
            const float b0 = p.b_scale_per_col ? p.b_scale[min(col, p.N - 1)] : b_s_uniform;
            const float b1 = p.b_scale_per_col ? p.b_scale[min(col + 1, p.N - 1)] : b_s_uniform;
            if (p.swap_ab) {                      // rows = channels (scale_b), columns = tokens (scale_a): same product order
              x0 = b0 * (a_s * x0);
              x1 = b1 * (a_s * x1);
            } else {
              x0 = a_s * (b0 * x0);               // ScaledEpilogue order: scale_a * (scale_b * acc)
              x1 = a_s * (b1 * x1);
            }
          }
          if (p.bias) {
            if (p.swap_ab) {
              const float bv = __bfloat162float(p.bias[min(row, p.M - 1)]);
              x0 += bv;
              x1 += bv;
            } else {
              x0 += __bfloat162float(p.bias[min(col, p.N - 1)]);
              x1 += __bfloat162float(p.bias[min(col + 1, p.N - 1)]);
            }
          }
          o[j >> 1] = pack_bf16x2(x0, x1);
        }
        if (p.swap_ab) {
          // transposed store: this thread owns output channel `row`, registers hold tokens col0 .. col0 + 31; a warp writes
          // 32 consecutive channels (64 bytes) of one token row per instruction
          if (row < p.M) {
            const __nv_bfloat16* ob = reinterpret_cast<const __nv_bfloat16*>(o);
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) p.c[(int64_t)(col0 + j) * p.ldc + row] = ob[j];
          }
          continue;
        }
        if (p.use_tma_store && (kBlockN / 32) % 2 == 1 && c == kBlockN / 32 - 1) {
          // an odd number of 32-column chunks (BLOCK_N 224): the last one has no partner for a 64-column TMA box; four
          // 16-byte stores per row instead (C rows are 16-byte aligned in this mode)
          if (row < p.M && col0 < p.N) {
            uint4* dst = reinterpret_cast<uint4*>(p.c + (int64_t)row * p.ldc + col0);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (col0 + 8 * j < p.N) dst[j] = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
          }
        } else if (p.use_tma_store) {
          // staging tile [32 rows][64 cols] bf16, row = 128 bytes, 16-byte chunk j stored at (j ^ (row & 7))
          uint8_t* st = my_stage + sbuf * 4096;
          if ((c & 1) == 0) tma_store_wait_read<1>();   // the buffer we are about to overwrite was stored 2 groups ago
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int chunk = (c & 1) * 4 + j;
            *reinterpret_cast<uint4*>(st + lane * 128 + ((chunk ^ (lane & 7)) << 4)) =
                make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
          }
          if ((c & 1) == 1) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0 && row0 < p.M && col0 - 32 < p.N) {
              tma_store_2d(&tmap_c, st, col0 - 32, row0);
              tma_store_commit();
            } else if (lane == 0) {
              tma_store_commit();   // keep the group count in step
            }
            sbuf ^= 1;
          }
        } else if (row < p.M && col0 < p.N) {
          __nv_bfloat16* dst = p.c + (int64_t)row * p.ldc + col0;
          const __nv_bfloat16* ob = reinterpret_cast<const __nv_bfloat16*>(o);
          for (int j = 0; j < 32 && col0 + j < p.N; ++j) dst[j] = ob[j];
        }
      }
      }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        if (kPair) mbar_arrive_cluster(tmem_empty_leader + as * 8);
        else mbar_arrive(tmem_empty + as);
      }
      if (++as == 2) { as = 0; aph ^= 1; }
    }
    if (p.use_tma_store && lane == 0) tma_store_wait_all();
  } else if (kind_is_wq(kKind) && warp >= 2 + kNumEpiWarps) {
    // ======================= W4 / W8 converters (warps 6..13, pairs: 6..21 in two sets) =======================
    // converter warp cw handles row tiles cw, cw+8, ... of the stage; a lane's 16 bytes of packed data are
    // rows (g, g+8) x k in [16t, 16t+16) of its row tile  ->  four 16-byte chunks of the swizzled bf16 tile.
    const int cw = (warp - (2 + kNumEpiWarps)) % Cfg::kConvPerStage;     // row-tile lane of this warp within its set
    const int cset = (warp - (2 + kNumEpiWarps)) / Cfg::kConvPerStage;    // which alternate k blocks this warp converts
    const int g = lane >> 2, t = lane & 3;
    int it = 0;                                  // k blocks since kernel start (every set walks all of them)
    for (int tile = unit0; tile < num_tiles; tile += unit_stride) {
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        if (Cfg::kConvSets > 1 && (it % Cfg::kConvSets) != cset) continue;
        const int s = it % kStages, bs = it % Cfg::kBRing;
        const uint32_t ph = (it / kStages) & 1, bph = (it / Cfg::kBRing) & 1;
        mbar_wait(packed_bar + s, ph);          // packed tile + meta landed
        mbar_wait(bempty_bar + bs, bph ^ 1);    // the MMA has finished reading this B buffer
        // explicit shared-space addresses: through generic pointers these compile to LD.E / ST.E (address-space
        // resolution per access) instead of LDS / STS
        const uint32_t pk = smem_u32(stage_packed(s));
        const uint32_t mt = smem_u32(stage_meta(s));
        const uint32_t bt = smem_u32(b_ring(bs));
#pragma unroll
        for (int r = cw; r < Cfg::kCtaN / 16; r += Cfg::kConvPerStage) {
          if (p.debug_skip_convert) break;
          uint32_t lo[8], hi[8];   // row g / row g+8: 16 bf16 = 8 packed registers, k ascending
          const uint32_t m0 = lds_32(mt + (r * 16 + g) * 4), m1 = lds_32(mt + (r * 16 + g + 8) * 4);
          if constexpr (kKind == kKindW8) {
            // int8 tile: 32 bytes per lane (row g then row g+8), meta = bf16 scale | zero << 16 (common.cuh w8_dequant_word)
            const uint4 wa = lds_128(pk + r * 1024 + lane * 32);
            const uint4 wb = lds_128(pk + r * 1024 + lane * 32 + 16);
            uint32_t s0, s1;
            float nz0, nz1;
            w8_meta(m0, s0, nz0);
            w8_meta(m1, s1, nz1);
            const uint32_t* av = &wa.x;
            const uint32_t* bv = &wb.x;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              w8_dequant_word(av[j], nz0, s0, lo[2 * j], lo[2 * j + 1]);
              w8_dequant_word(bv[j], nz1, s1, hi[2 * j], hi[2 * j + 1]);
            }
          } else {
          const uint4 wq = lds_128(pk + r * 512 + lane * 16);
          const uint32_t s0 = __byte_perm(m0, 0, 0x1010), z0 = __byte_perm(m0, 0, 0x3232);
          const uint32_t s1 = __byte_perm(m1, 0, 0x1010), z1 = __byte_perm(m1, 0, 0x3232);
          const uint32_t* wv = &wq.x;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t w = wv[j];
            uint32_t q0, q1, q2, q3;
            asm("lop3.b32 %0, %1, 0x000f000f, 0x43004300, 0xea;" : "=r"(q0) : "r"(w));
            asm("lop3.b32 %0, %1, 0x000f000f, 0x43004300, 0xea;" : "=r"(q1) : "r"(w >> 4));
            asm("lop3.b32 %0, %1, 0x000f000f, 0x43004300, 0xea;" : "=r"(q2) : "r"(w >> 8));
            asm("lop3.b32 %0, %1, 0x000f000f, 0x43004300, 0xea;" : "=r"(q3) : "r"(w >> 12));
            uint32_t d;
            asm("sub.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(q0), "r"(z0));
            asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(lo[2 * j]) : "r"(d), "r"(s0));
            asm("sub.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(q2), "r"(z0));
            asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(lo[2 * j + 1]) : "r"(d), "r"(s0));
            asm("sub.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(q1), "r"(z1));
            asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(hi[2 * j]) : "r"(d), "r"(s1));
            asm("sub.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(q3), "r"(z1));
            asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(hi[2 * j + 1]) : "r"(d), "r"(s1));
          }
          }
          // swizzled K-major tile: row n at n*128 bytes, 16-byte chunk c stored at chunk (c ^ (n & 7))
          const int n_lo = r * 16 + g, n_hi = n_lo + 8;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int c = 2 * t + h;
            sts_128(bt + n_lo * 128 + ((c ^ (n_lo & 7)) << 4), make_uint4(lo[4 * h], lo[4 * h + 1], lo[4 * h + 2], lo[4 * h + 3]));
            sts_128(bt + n_hi * 128 + ((c ^ (n_hi & 7)) << 4), make_uint4(hi[4 * h], hi[4 * h + 1], hi[4 * h + 2], hi[4 * h + 3]));
          }
        }
        fence_proxy_async_smem();     // generic-proxy stores -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) {
          if (kPair) mbar_arrive_cluster(bready_leader + bs * 8);   // the leader's MMA reads this CTA's half through the pair
          else mbar_arrive(bready_bar + bs);
        }
      }
    }
  }

  tc_fence_before_sync();
  if (kPair) cluster_sync_all();   // no CTA leaves (or frees TMEM) while its peer may still signal its barriers / read its B half
  else __syncthreads();
  if (warp == 1) {
    if (kPair) tmem_dealloc_cg2(tmem_base, Cfg::kTmemCols);
    else tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <int kKind, int kBlockN, int kMT = 1, int kCG = 1>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, GemmParams p, cudaStream_t stream,
                       const CUtensorMap* tm = nullptr) {
  using Cfg = GemmCfg<kKind, kBlockN, kMT, kCG>;
  auto kern = gemm_tcgen05_kernel<kKind, kBlockN, kMT, kCG>;
  static bool attr_done = false;
  static int max_ctas = 148;     // persistent CTAs: one per SM; CTA pairs: 2 x the pairs the device co-schedules
  if (!attr_done) {
    XB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_ctas, cudaDevAttrMultiProcessorCount, dev);
    if (kCG == 2) {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(max_ctas & ~1);
      cfg.blockDim = dim3(Cfg::kThreads);
      cfg.dynamicSmemBytes = Cfg::kSmemBytes;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 2;
      at[0].val.clusterDim.y = 1;
      at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) == cudaSuccess && n > 0) max_ctas = 2 * n;
      else max_ctas &= ~1;
      cudaGetLastError();
    }
    attr_done = true;
  }
  const int units = ((p.M + Cfg::kCtaM * kCG - 1) / (Cfg::kCtaM * kCG)) * ((p.N + kBlockN - 1) / kBlockN) * (p.split_k > 1 ? p.split_k : 1);
  if (p.split_k < 1) p.split_k = 1;
  const int want = units * kCG;
  dim3 grid(want < max_ctas ? want : max_ctas), block(Cfg::kThreads);
  CUtensorMap tc = ta;   // placeholder when the direct-store epilogue is used
  p.use_tma_store = !p.swap_ab && (p.ldc % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.c) & 15) == 0);
  static const int dbg_skip = [] { const char* e = getenv("XB_GEMM_DEBUG_SKIP_CONVERT"); return e ? atoi(e) : 0; }();
  p.debug_skip_convert = dbg_skip;
  if (p.use_tma_store && make_tmap_2d(&tc, p.c, p.M, p.N, (uint64_t)p.ldc * 2, 32, 64, 2)) return 1;
  XB_CUDA_OK(launch_cluster(kern, grid, block, (size_t)Cfg::kSmemBytes, stream, true, kCG, ta, tb, tc, tm ? *tm : ta, p));
  return 0;
}

}  // namespace tc
}  // namespace xb

using namespace xb;
using namespace xb::tc;

// pick BLOCK_N so that small problems still spread over the SMs (the reference buckets by M the same way:
// scaled_mm_sm100_fp8_dispatch.cuh:148-287)
static int pick_block_n(int M, int N) {
  const int m_blocks = (M + kBlockM - 1) / kBlockM;
  static const int force = [] { const char* e = getenv("XB_GEMM_BN"); return e ? atoi(e) : 0; }();
  if (force == 64 || force == 128 || force == 256) return force;
  if ((int64_t)m_blocks * ((N + 255) / 256) >= 140) return 256;   // UMMA N=256: A re-read from smem half as often
  if ((int64_t)m_blocks * ((N + 127) / 128) >= 120) return 128;
  return 64;
}

// CTA-pair (cta_group::2) kernel: 256 x 256 tiles, each CTA stages half of B, so the shared-memory traffic per MMA flop
// drops by a third (bf16 / fp8) and the converter work per flop halves (W4 / W8).  Needs enough 256-row tiles to fill
// the 74 pairs.  Mode: xb_set_gemm_cta_pair() / XB_GEMM_CG.
static std::atomic<int> g_cta_pair_mode{-1};   // -1: not set -> XB_GEMM_CG, else 0 (auto)
static bool use_cta_pair(int M, int N, int bn, bool weight_only = false) {
  int mode = g_cta_pair_mode.load(std::memory_order_relaxed);
  if (mode < 0) {
    const char* e = getenv("XB_GEMM_CG");
    mode = e ? atoi(e) : 0;
    g_cta_pair_mode.store(mode, std::memory_order_relaxed);
  }
  // automatic = pairs.  Measured on B200 at M = 8192 (tools/gemm_sweep.py, profiles/r02d_gemm_sweep.md): bf16 +8..18 %
  // (1.35-1.50 PF/s, 0.92-1.02 x cuBLASLt), fp8 +8..17 % (2.59-2.78 PF/s, 0.99-1.09 x torch._scaled_mm), W4A16 +13..18 %
  // (0.90-1.20 PF/s) over the single-CTA kernels
  (void)weight_only;
  if (mode == 0) mode = 2;
  if (mode == 1 || bn != 256 || N % 256 != 0) return false;
  const int64_t units = (int64_t)((M + 255) / 256) * (N / 256);
  if (mode == 3) return M > 128;                   // every shape that can form a pair tile (tests)
  return M >= 512 && units >= 74;
}

static std::atomic<int> g_fp8_swap_max_m{-1};
static int fp8_swap_max_m() {
  int v = g_fp8_swap_max_m.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("XB_FP8_SWAP_MAX_M");
    // default 64: measured on B200 at M = 32 (tools/fp8_swap_probe.py, profiles/r02f_fp8_swap.md) swap-AB is 1.0-1.8x the
    // token-major tile on the Llama-3-70B projections (qkv 43 -> 24 us unsharded, down TP8 shard 16.4 -> 12.7 us)
    v = e ? atoi(e) : 64;
    if (v < 0 || v > 64) v = 64;
    g_fp8_swap_max_m.store(v, std::memory_order_relaxed);
  }
  return v;
}
// BLOCK_N of a pair tile: 256, or 224 when that fills the last wave of the 74 pairs better (N = 3584: 14 x 32 = 448 tiles are
// 6.05 waves = 86 % of 7 waves; 16 x 32 = 512 tiles of 224 columns are 6.92 waves = 99 %, at ~3 % less operand reuse)
static int pair_block_n(int M, int N) {
  static const int force = [] { const char* e = getenv("XB_GEMM_PAIR_BN"); return e ? atoi(e) : 0; }();
  if (force == 256 || N % 224 != 0) return 256;
  if (force == 224) return 224;
  if (N % 256 != 0) return 224;
  const int64_t mp = (M + 255) / 256;
  auto eff = [&](int bn) {
    const int64_t units = mp * (N / bn);
    const int64_t waves = (units + 73) / 74;
    return (double)units / (double)(waves * 74) * (bn == 224 ? 0.97 : 1.0);
  };
  return eff(224) > eff(256) ? 224 : 256;
}

// split-K workspace of the swap-AB FP8 GEMM (caller-owned device memory; every split-K launch uses it, so launches that share
// it must be stream-ordered - the decode step's are).  Layout: 4 KB of tickets (zeroed here, self-resetting afterwards), then
// the fp32 partials.
static std::atomic<void*> g_splitk_ws{nullptr};
static std::atomic<size_t> g_splitk_bytes{0};
static std::atomic<int> g_splitk_device{-1};     // the device the workspace lives on: GEMMs on another device run unsplit
static std::atomic<int> g_splitk_max{-1};
constexpr size_t kSplitKTicketBytes = 4096;
static int splitk_max() {
  int v = g_splitk_max.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("XB_FP8_SPLITK_MAX");
    v = e ? atoi(e) : 8;
    if (v < 1 || v > 32) v = 8;
    g_splitk_max.store(v, std::memory_order_relaxed);
  }
  return v;
}
// k ranges for `tiles` output tiles over `num_kb` k blocks: fill the SMs once (tiles * split <= sm_count), keep >= 12 k blocks
// per range and no empty range.  Measured on B200 at M = 32 (tools/fp8_splitk_probe.py, profiles/r02i_fp8_splitk.md): ranges of
// 4 / 8 k blocks LOSE (o_proj TP8 shard, K = 1024: 7.3 -> 8.7 us; TP4, K = 2048: 9.9 -> 10.2 us - ring fill + the ticket round
// trip outweigh the extra CTAs), 14 and more win (down TP8 13.3 -> 12.1 us, qkv TP8 22.5 -> 11.8 us, down unsharded 70 -> 46 us)
static int split_k_for(int tiles, int num_kb, int sms) {
  int split = sms / tiles;
  if (split > splitk_max()) split = splitk_max();
  if (split > num_kb / 12) split = num_kb / 12;
  if (split < 2) return 1;
  const int kb_per = (num_kb + split - 1) / split;
  split = (num_kb + kb_per - 1) / kb_per;
  return split < 2 ? 1 : split;
}
static int pick_split_k(int tiles, int num_kb, int tokens_bn, size_t* need_bytes) {
  *need_bytes = 0;
  if (g_splitk_ws.load(std::memory_order_relaxed) == nullptr) return 1;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev != g_splitk_device.load(std::memory_order_relaxed)) {
    cudaGetLastError();
    return 1;
  }
  static int sms = [] { int d = 0, n = 148; cudaGetDevice(&d); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d); return n; }();
  int split = split_k_for(tiles, num_kb, sms);
  if (split < 2) return 1;
  const size_t need = kSplitKTicketBytes + (size_t)tiles * split * tokens_bn * kBlockM * sizeof(float);
  if ((size_t)tiles * sizeof(unsigned) > kSplitKTicketBytes || need > g_splitk_bytes.load(std::memory_order_relaxed)) return 1;
  *need_bytes = need;
  return split;
}

// host-only query (no device needed): the number of k ranges xb_gemm_fp8_scaled would use for an [M, K] x [N, K]^T product on a
// device with `sm_count` SMs once a workspace is registered; 1 = unsplit.  For capacity planning and the CPU tests.
extern "C" int xb_gemm_fp8_split_k(int M, int N, int K, int sm_count) {
  if (M <= 0 || N < 128 || K <= 0 || sm_count <= 0 || M > 64 || M > fp8_swap_max_m()) return 1;
  return split_k_for((N + kBlockM - 1) / kBlockM, (K + 127) / 128, sm_count);
}

// host-only (no device needed): which kernel variant the tcgen05 GEMM entry points pick for a shape, as text, e.g.
// "bf16 pair 256x224", "fp8 swap-AB bn=32 split_k=5", "w4 single bn=128".  kind: 0 bf16 | 1 fp8 | 2 w4a16 | 3 w8a16.  Mirrors
// the decision order of xb_gemm_bf16 / xb_gemm_fp8_scaled / xb_gemm_w4a16 / xb_gemm_w8a16 below (keep the two in step); split_k is
// what a registered workspace would allow on a device with sm_count SMs.  Returns the length written, -1 on a bad kind.
extern "C" int xb_gemm_describe(int kind, int M, int N, int K, int sm_count, char* buf, int buf_len) {
  static const char* names[] = {"bf16", "fp8", "w4", "w8"};
  if (kind < 0 || kind > 3 || buf == nullptr || buf_len <= 0) return -1;
  if (kind == 1 && M <= fp8_swap_max_m() && M <= 64 && N >= 128)
    return snprintf(buf, (size_t)buf_len, "fp8 swap-AB bn=%d split_k=%d", M <= 32 ? 32 : 64, xb_gemm_fp8_split_k(M, N, K, sm_count));
  int bn = pick_block_n(M, N);
  bool pair;
  int pbn = 0;
  if (kind == 3) {
    pair = use_cta_pair(M, N, bn, true);
    if (bn == 256 && !pair) bn = 128;
    if (bn == 128 && N % 128 != 0) bn = 64;
    pbn = pair ? 256 : 0;
  } else {
    if (kind == 2) {
      if (bn == 256 && N % 256 != 0) bn = 128;
      if (bn == 128 && N % 128 != 0) bn = 64;
    }
    pair = use_cta_pair(M, N, bn, kind == 2);
    pbn = pair ? pair_block_n(M, N) : 0;
  }
  if (pair) return snprintf(buf, (size_t)buf_len, "%s pair 256x%d", names[kind], pbn);
  return snprintf(buf, (size_t)buf_len, "%s single bn=%d", names[kind], bn);
}

extern "C" size_t xb_gemm_splitk_workspace_bytes(void) {
  // one wave of 148 CTAs x the largest token tile (64) is the most the heuristic ever asks for
  return kSplitKTicketBytes + (size_t)148 * 64 * kBlockM * sizeof(float);
}

extern "C" int xb_set_gemm_splitk_workspace(void* ws, size_t bytes) {
  if (ws != nullptr && (bytes <= kSplitKTicketBytes || (reinterpret_cast<uintptr_t>(ws) & 15) != 0)) {
    set_error("set_gemm_splitk_workspace: need a 16-byte aligned buffer of more than %zu bytes", kSplitKTicketBytes);
    return -1;
  }
  int dev = -1;
  if (ws != nullptr) {
    XB_CUDA_OK(cudaMemset(ws, 0, kSplitKTicketBytes));
    cudaPointerAttributes attr{};
    if (cudaPointerGetAttributes(&attr, ws) == cudaSuccess) dev = attr.device;
    else { cudaGetLastError(); cudaGetDevice(&dev); }
  }
  g_splitk_device.store(dev, std::memory_order_relaxed);
  g_splitk_ws.store(ws, std::memory_order_relaxed);
  g_splitk_bytes.store(ws ? bytes : 0, std::memory_order_relaxed);
  return 0;
}

extern "C" int xb_set_fp8_splitk_max(int max_split) {
  if (max_split < 1 || max_split > 32) {
    set_error("set_fp8_splitk_max: %d out of range (1 = never split .. 32)", max_split);
    return -1;
  }
  const int old = splitk_max();
  g_splitk_max.store(max_split, std::memory_order_relaxed);
  return old;
}

extern "C" int xb_set_fp8_swap_max_m(int max_m) {
  if (max_m < 0 || max_m > 64) {
    set_error("set_fp8_swap_max_m: %d out of range (0 = never swap .. 64)", max_m);
    return -1;
  }
  const int old = fp8_swap_max_m();
  g_fp8_swap_max_m.store(max_m, std::memory_order_relaxed);
  return old;
}

extern "C" int xb_set_gemm_cta_pair(int mode) {
  if (mode < 0 || mode > 3) {
    set_error("set_gemm_cta_pair: mode %d (0 auto | 1 single | 2 pairs | 3 pairs always)", mode);
    return -1;
  }
  const int old = g_cta_pair_mode.exchange(mode, std::memory_order_relaxed);
  return old < 0 ? 0 : old;
}

extern "C" int xb_gemm_bf16(void* c, int64_t ldc, const void* a, int64_t lda, const void* b, const void* bias, int M, int N,
                            int K, xb_stream_t stream) {
  if (M == 0 || N == 0) return 0;
  XB_CHECK(K % 8 == 0 && lda % 8 == 0, "gemm_bf16: K=%d and lda=%lld must be multiples of 8 (16-byte rows for TMA)", K,
           (long long)lda);
  XB_CHECK((reinterpret_cast<uintptr_t>(a) & 15) == 0 && (reinterpret_cast<uintptr_t>(b) & 15) == 0, "gemm_bf16: a/b alignment");
  GemmParams p{};
  p.c = reinterpret_cast<__nv_bfloat16*>(c);
  p.ldc = ldc;
  p.bias = reinterpret_cast<const __nv_bfloat16*>(bias);
  p.M = M; p.N = N; p.K = K;
  const int bn = pick_block_n(M, N);
  const bool pair = use_cta_pair(M, N, bn);
  const int pbn = pair ? pair_block_n(M, N) : 0;
  CUtensorMap ta, tb;
  if (make_tmap_2d(&ta, a, M, K, (uint64_t)lda * 2, kBlockM, 64, 2)) return 1;
  if (make_tmap_2d(&tb, b, N, K, (uint64_t)K * 2, pair ? pbn / 2 : bn, 64, 2)) return 1;
  if (pair && pbn == 224) return launch_gemm<kKindBF16, 224, 1, 2>(ta, tb, p, (cudaStream_t)stream);
  if (pair) return launch_gemm<kKindBF16, 256, 1, 2>(ta, tb, p, (cudaStream_t)stream);
  if (bn == 256) return launch_gemm<kKindBF16, 256>(ta, tb, p, (cudaStream_t)stream);
  return bn == 128 ? launch_gemm<kKindBF16, 128>(ta, tb, p, (cudaStream_t)stream)
                   : launch_gemm<kKindBF16, 64>(ta, tb, p, (cudaStream_t)stream);
}

extern "C" int xb_gemm_fp8_scaled(void* c, int64_t ldc, const void* a, int64_t lda, const void* b, const float* a_scale,
                                  int a_scale_numel, const float* b_scale, int b_scale_numel, const void* bias, int M, int N,
                                  int K, xb_stream_t stream) {
  if (M == 0 || N == 0) return 0;
  // same argument checks as cutlass_scaled_mm (scaled_mm_entry.cu:62-85)
  XB_CHECK(K % 16 == 0 && lda % 16 == 0, "cutlass_scaled_mm: K=%d / lda=%lld must be multiples of 16", K, (long long)lda);
  XB_CHECK(ldc % 8 == 0, "cutlass_scaled_mm: c.stride(0) %% 16 bytes != 0");
  XB_CHECK(a_scale_numel == 1 || a_scale_numel == M, "cutlass_scaled_mm: a_scales must have numel 1 or M");
  XB_CHECK(b_scale_numel == 1 || b_scale_numel == N, "cutlass_scaled_mm: b_scales must have numel 1 or N");
  GemmParams p{};
  p.c = reinterpret_cast<__nv_bfloat16*>(c);
  p.ldc = ldc;
  p.bias = reinterpret_cast<const __nv_bfloat16*>(bias);
  p.a_scale = a_scale; p.b_scale = b_scale;
  p.a_scale_per_row = a_scale_numel > 1; p.b_scale_per_col = b_scale_numel > 1;
  p.M = M; p.N = N; p.K = K;
  // decode-sized M: swap-AB (see GemmParams::swap_ab).  Measured on B200 (tools/fp8_swap_probe.py) against the token-major
  // tile and the mma.sync streaming kernel.  XB_FP8_SWAP_MAX_M: largest M that swaps (0 disables).
  if (M <= fp8_swap_max_m() && M <= 64 && N >= 128) {
    p.swap_ab = 1;
    p.M = N; p.N = M;
    p.a_scale = b_scale; p.b_scale = a_scale;          // per-row = channel scales, per-column = token scales
    p.a_scale_per_row = b_scale_numel > 1; p.b_scale_per_col = a_scale_numel > 1;
    {
      size_t need = 0;
      const int split = pick_split_k((N + kBlockM - 1) / kBlockM, (K + 127) / 128, M <= 32 ? 32 : 64, &need);
      if (split > 1) {
        uint8_t* ws = static_cast<uint8_t*>(g_splitk_ws.load(std::memory_order_relaxed));
        p.split_k = split;
        p.splitk_tickets = reinterpret_cast<unsigned int*>(ws);
        p.splitk_ws = reinterpret_cast<float*>(ws + kSplitKTicketBytes);
      }
    }
    CUtensorMap tw, tx;
    if (make_tmap_2d(&tw, b, N, K, (uint64_t)K, kBlockM, 128, 1)) return 1;
    if (M <= 32) {
      if (make_tmap_2d(&tx, a, M, K, (uint64_t)lda, 32, 128, 1)) return 1;
      return launch_gemm<kKindFP8, 32>(tw, tx, p, (cudaStream_t)stream);
    }
    if (make_tmap_2d(&tx, a, M, K, (uint64_t)lda, 64, 128, 1)) return 1;
    return launch_gemm<kKindFP8, 64>(tw, tx, p, (cudaStream_t)stream);
  }
  const int bn = pick_block_n(M, N);
  const bool pair = use_cta_pair(M, N, bn);
  const int pbn = pair ? pair_block_n(M, N) : 0;
  CUtensorMap ta, tb;
  if (make_tmap_2d(&ta, a, M, K, (uint64_t)lda, kBlockM, 128, 1)) return 1;
  if (make_tmap_2d(&tb, b, N, K, (uint64_t)K, pair ? pbn / 2 : bn, 128, 1)) return 1;
  if (pair && pbn == 224) return launch_gemm<kKindFP8, 224, 1, 2>(ta, tb, p, (cudaStream_t)stream);
  if (pair) return launch_gemm<kKindFP8, 256, 1, 2>(ta, tb, p, (cudaStream_t)stream);
  if (bn == 256) return launch_gemm<kKindFP8, 256>(ta, tb, p, (cudaStream_t)stream);
  return bn == 128 ? launch_gemm<kKindFP8, 128>(ta, tb, p, (cudaStream_t)stream)
                   : launch_gemm<kKindFP8, 64>(ta, tb, p, (cudaStream_t)stream);
}

extern "C" int xb_gemm_w4a16(void* c, int64_t ldc, const void* a, int64_t lda, const uint32_t* qweight, const uint32_t* meta,
                             const void* bias, int M, int N, int K, int group_size, xb_stream_t stream) {
  if (M == 0 || N == 0) return 0;
  XB_CHECK(N % 64 == 0 && K % 64 == 0, "gemm_w4a16: N=%d and K=%d must be multiples of 64", N, K);
  XB_CHECK(group_size >= 64 && K % group_size == 0, "gemm_w4a16: bad group_size %d", group_size);
  const int tpg = group_size / 64;
  XB_CHECK((tpg & (tpg - 1)) == 0, "gemm_w4a16: group_size/64 must be a power of two");
  XB_CHECK(lda % 8 == 0 && (reinterpret_cast<uintptr_t>(a) & 15) == 0, "gemm_w4a16: a alignment");
  GemmParams p{};
  p.c = reinterpret_cast<__nv_bfloat16*>(c);
  p.ldc = ldc;
  p.bias = reinterpret_cast<const __nv_bfloat16*>(bias);
  p.qweight = reinterpret_cast<const uint4*>(qweight);
  p.meta = meta;
  p.gshift = 0;
  while ((1 << p.gshift) < tpg) ++p.gshift;
  p.M = M; p.N = N; p.K = K;
  int bn = pick_block_n(M, N);
  if (bn == 256 && N % 256 != 0) bn = 128;
  if (bn == 128 && N % 128 != 0) bn = 64;
  // XB_GEMM_W4_MT=2: two 128-row M tiles per CTA share every dequantised B stage (BLOCK_N 128), halving the converter
  // work per MMA flop.  Measured on B200 it is ~5 % SLOWER than the single-tile BLOCK_N 256 kernel (1.05 vs 1.10 PF/s):
  // both move ~115-120 KB of shared memory per 512 MMA cycles, i.e. the kernel is shared-memory-bandwidth bound, not
  // converter bound; the fix is cta_group::2 (each CTA converts and holds half of B).  Kept as an experiment switch.
  static const int force_mt = [] { const char* e = getenv("XB_GEMM_W4_MT"); return e ? atoi(e) : 0; }();
  const bool two_m = force_mt == 2 && M >= 256 && N % 128 == 0;
  if (two_m) bn = 128;
  const bool pair = !two_m && use_cta_pair(M, N, bn, true);
  const int pbn = pair ? pair_block_n(M, N) : 0;
  const int cta_n = pair ? pbn / 2 : bn;     // B rows one CTA stages and dequantises
  CUtensorMap ta, tb, tm;
  if (make_tmap_2d(&ta, a, M, K, (uint64_t)lda * 2, two_m ? 256 : kBlockM, 64, 2)) return 1;
  const uint64_t ktiles = K / 64;
  if (make_tmap_2d_raw(&tb, qweight, CU_TENSOR_MAP_DATA_TYPE_UINT32, N / 16, ktiles * 128, ktiles * 512, cta_n / 16, 128,
                       CU_TENSOR_MAP_SWIZZLE_NONE))
    return 1;
  if (make_tmap_2d_raw(&tm, meta, CU_TENSOR_MAP_DATA_TYPE_UINT32, K / group_size, N, (uint64_t)N * 4, 1, cta_n,
                       CU_TENSOR_MAP_SWIZZLE_NONE))
    return 1;
  if (pair && pbn == 224) return launch_gemm<kKindW4, 224, 1, 2>(ta, tb, p, (cudaStream_t)stream, &tm);
  if (pair) return launch_gemm<kKindW4, 256, 1, 2>(ta, tb, p, (cudaStream_t)stream, &tm);
  if (two_m) return launch_gemm<kKindW4, 128, 2>(ta, tb, p, (cudaStream_t)stream, &tm);
  if (bn == 256) return launch_gemm<kKindW4, 256>(ta, tb, p, (cudaStream_t)stream, &tm);
  return bn == 128 ? launch_gemm<kKindW4, 128>(ta, tb, p, (cudaStream_t)stream, &tm)
                   : launch_gemm<kKindW4, 64>(ta, tb, p, (cudaStream_t)stream, &tm);
}

// kind W8: B is the tile-packed int8 weight of linear_q8_small_m.cu ([N/16][K/64][32 lanes][8 words]); same pipeline as W4
// with a 1 KB packed row tile per k64 step and the int8 converter (w8_dequant_word).
extern "C" int xb_gemm_w8a16(void* c, int64_t ldc, const void* a, int64_t lda, const uint32_t* qweight, const uint32_t* meta,
                             const void* bias, int M, int N, int K, int group_size, xb_stream_t stream) {
  if (M == 0 || N == 0) return 0;
  XB_CHECK(N % 64 == 0 && K % 64 == 0, "gemm_w8a16: N=%d and K=%d must be multiples of 64", N, K);
  XB_CHECK(group_size >= 64 && K % group_size == 0, "gemm_w8a16: bad group_size %d", group_size);
  const int tpg = group_size / 64;
  XB_CHECK((tpg & (tpg - 1)) == 0, "gemm_w8a16: group_size/64 must be a power of two");
  XB_CHECK(lda % 8 == 0 && (reinterpret_cast<uintptr_t>(a) & 15) == 0, "gemm_w8a16: a alignment");
  GemmParams p{};
  p.c = reinterpret_cast<__nv_bfloat16*>(c);
  p.ldc = ldc;
  p.bias = reinterpret_cast<const __nv_bfloat16*>(bias);
  p.qweight = reinterpret_cast<const uint4*>(qweight);
  p.meta = meta;
  p.gshift = 0;
  while ((1 << p.gshift) < tpg) ++p.gshift;
  p.M = M; p.N = N; p.K = K;
  int bn = pick_block_n(M, N);
  const bool pair = use_cta_pair(M, N, bn, true);   // a pair stages 2 x 128 rows of int8: the single-CTA 256-row tile does not fit
  if (bn == 256 && !pair) bn = 128;           // 256 rows of int8 + the bf16 ring leave too few TMA stages
  if (bn == 128 && N % 128 != 0) bn = 64;
  const int cta_n = pair ? bn / 2 : bn;
  CUtensorMap ta, tb, tm;
  if (make_tmap_2d(&ta, a, M, K, (uint64_t)lda * 2, kBlockM, 64, 2)) return 1;
  const uint64_t ktiles = K / 64;
  if (make_tmap_2d_raw(&tb, qweight, CU_TENSOR_MAP_DATA_TYPE_UINT32, N / 16, ktiles * 256, ktiles * 1024, cta_n / 16, 256,
                       CU_TENSOR_MAP_SWIZZLE_NONE))
    return 1;
  if (make_tmap_2d_raw(&tm, meta, CU_TENSOR_MAP_DATA_TYPE_UINT32, K / group_size, N, (uint64_t)N * 4, 1, cta_n,
                       CU_TENSOR_MAP_SWIZZLE_NONE))
    return 1;
  if (pair) return launch_gemm<kKindW8, 256, 1, 2>(ta, tb, p, (cudaStream_t)stream, &tm);
  return bn == 128 ? launch_gemm<kKindW8, 128>(ta, tb, p, (cudaStream_t)stream, &tm)
                   : launch_gemm<kKindW8, 64>(ta, tb, p, (cudaStream_t)stream, &tm);
}
