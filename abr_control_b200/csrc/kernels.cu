// kernels.cu — sm_100a kernels of the batched arm engine, compiled once per joint count (-DABRB_N=<n>).
//
// Mapping: ONE JOINT STATE PER THREAD.  The per-state arithmetic (abrb_math.cuh / abrb_rbd.cuh /
// abrb_osc.cuh) is a straight-line, fully unrolled sequence on registers; chain constants come from the
// kernel-parameter constant bank (uniform across the warp).  Loads of the (B, n) state arrays are
// contiguous per warp; every output array is staged per warp in shared memory and written back with
// fully coalesced 16-byte stores, so HBM sees whole 128-byte lines only.
#include <cuda_runtime.h>
#include <cstdlib>
#include <utility>

#include <atomic>

#include "abrb_coop.cuh"
#include "abrb_launch.hpp"
#include "abrb_osc.cuh"
#include "abrb_rbd.cuh"

#ifndef ABRB_N
#error "compile with -DABRB_N=<joint count>"
#endif

namespace abrb {

namespace {

constexpr int kBlock = 128;
// CTAs per SM the register allocator must allow.  Measured on B200 (tools/kbench.py, B = 65536): the fp64 kernels
// are fastest with the full 255 registers (2 CTAs/SM; capping them spills and costs 20-50 %), the fp32 kernels with
// 4 CTAs/SM (128 registers; the 6-DOF OSC kernel goes from 222 to 129 us).
#ifndef ABRB_KSMEM_F32
#define ABRB_KSMEM_F32 0  // 1: shared-memory kinematic scratch for the fp32 kernels too
#endif
template <typename T>
struct MinBlocks {
#ifndef ABRB_MINBLOCKS_F32
#define ABRB_MINBLOCKS_F32 4
#endif
#ifndef ABRB_MINBLOCKS_F64
#define ABRB_MINBLOCKS_F64 2
#endif
  static constexpr int value = sizeof(T) == 8 ? ABRB_MINBLOCKS_F64 : ABRB_MINBLOCKS_F32;
};
constexpr int kWarps = kBlock / 32;

template <int N>
struct MaxRecord {
  static constexpr int big = N * N > 6 * N ? N * N : 6 * N;
  static constexpr int value = big > 16 ? big : 16;
};

// Per-warp shared-memory region: holds the warp's kinematic scratch (slot-major, stride 32: lane i owns column i,
// conflict-free) while a state is being evaluated and is then re-used as the staging tile for the coalesced
// stores.  KSMEM selects the shared-memory scratch (fp64 kernels: keeps them under the register limit without
// local-memory spills); otherwise the scratch lives in registers and only the staging tile is needed.
template <typename T, int N, bool ORTHO, bool KSMEM>
struct KinSel;
template <typename T, int N, bool ORTHO>
struct KinSel<T, N, ORTHO, false> {
  typedef Kin<T, N, ORTHO, RegStore> type;
  static constexpr int kSlots = 0;
  static __device__ __forceinline__ void bind(type &, T *, int) {}
};
template <typename T, int N, bool ORTHO>
struct KinSel<T, N, ORTHO, true> {
  typedef Kin<T, N, ORTHO, StridedStore> type;
  static constexpr int kSlots = KinSlots<N, ORTHO>::kCount;
  static __device__ __forceinline__ void bind(type &k, T *warp_region, int lane) {
    k.s.base = warp_region + lane;
    k.s.stride = 32;
  }
};
template <int A, int B>
struct MaxI {
  static constexpr int value = A > B ? A : B;
};

// Write one LEN-element record per lane to `out[(warp_b0 + lane) * LEN + e]` through the warp's staging tile.
// The tile is element-major with a row pitch of 33 (`tile[e * 33 + lane]`): the per-lane writes are conflict-free
// (consecutive lanes -> consecutive words) and so are the reads of the linear copy-out (consecutive output
// elements -> pitch 33 -> distinct banks); the copy-out itself is perfectly coalesced (each warp instruction
// writes 32 consecutive elements = whole 128-byte lines).
constexpr int kPitch = 33;
// Records longer than kChunk elements go through the tile in slices of kChunk (the tile then needs only
// kPitch * kChunk elements per warp, which is what lets three CTAs of the fp64 kernels share an SM); each slice is a
// run of kChunk contiguous elements per record in global memory (>= 144 bytes for fp64).
constexpr int kChunk = 64;  // no slicing needed with 2 CTAs/SM (93 KB each); 18 would allow 3 CTAs/SM but measured slower
template <typename T, int LEN>
__device__ __forceinline__ void store_records(T *__restrict__ out, int64_t warp_b0, int nvalid, const T *rec,
                                              T *tile, int lane) {
  T *dst = out + warp_b0 * LEN;
#pragma unroll
  for (int c0 = 0; c0 < LEN; c0 += kChunk) {
    constexpr int kFull = kChunk;
    const int ch = LEN - c0 < kFull ? LEN - c0 : kFull;  // compile-time after unrolling
    __syncwarp();
#pragma unroll
    for (int e = 0; e < kChunk; ++e)
      if (e < ch) tile[e * kPitch + lane] = rec[c0 + e];
    __syncwarp();
    const int total = nvalid * ch;
#pragma unroll
    for (int it = 0; it < kChunk; ++it) {
      if (it < ch) {
        const int i = it * 32 + lane;
        const int r = i / ch, e = i - r * ch;
        if (i < total) dst[r * LEN + c0 + e] = tile[e * kPitch + r];
      }
    }
  }
}

// Output functor handed to rbd_state: stores each finished record immediately (keeps the live register set small)
template <typename T>
struct RbdSink {
  T *ptr[kOutCount];
  T *tile;
  int64_t warp_b0;
  int nvalid, lane;
  template <int LEN>
  __device__ __forceinline__ void put(int which, const T *rec) {
    if (ptr[which] != nullptr) store_records<T, LEN>(ptr[which], warp_b0, nvalid, rec, tile, lane);
  }
};

template <typename T>
struct RbdArgs {
  const T *q, *dq;
  T *Tx, *Tm, *R, *Tinv, *quat, *J, *dJ, *M, *g, *C;
  int64_t B;
  int frame;
  unsigned want;
  T xoff[3];
};

// shared memory per warp: [ kinematic scratch (KSMEM only): kSlots x 32 ][ staging tile: kPitch x max record ]
//                        [ exchange area of the cooperative pseudo-inverse (OSC kernels only): XCH x 32 ]
template <typename T, int N, bool ORTHO, bool KSMEM, int MAXREC, int XCH = 0>
struct WarpSmem {
  static constexpr int kKin = 32 * KinSel<T, N, ORTHO, KSMEM>::kSlots;
  static constexpr int kTile = kPitch * (MAXREC < kChunk ? MAXREC : kChunk);
  static constexpr int kXch = 32 * XCH;
  static constexpr int kElems = kKin + kTile + kXch;
};
template <typename T, int N, bool ORTHO, int KD, bool KSMEM>
struct OscSmem : WarpSmem<T, N, ORTHO, KSMEM, (N > 6 ? N : 6), CoopLayout<N, KD, !KSMEM>::kSlots> {};

// Programmatic dependent launch (see launch_pdl): let the stream's next kernel be scheduled as CTAs of this one retire,
// and wait until the previous kernel of the stream has completed and its memory is visible.  Both are no-ops for a
// launch without the attribute.
__device__ __forceinline__ void pdl_entry() {
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

template <typename T, int N, bool ORTHO, bool DYN, bool CMAT, bool XTRA, bool KSMEM>
__global__ void __launch_bounds__(kBlock, MinBlocks<T>::value)
rbd_kernel(const __grid_constant__ ChainK<T, N> P, const __grid_constant__ RbdArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  typedef KinSel<T, N, ORTHO, KSMEM> KS;
  typedef WarpSmem<T, N, ORTHO, KSMEM, MaxRecord<N>::value> WS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T *region = reinterpret_cast<T *>(smem_raw) + warp * WS::kElems;
  typename KS::type K;
  KS::bind(K, region, lane);
  RbdSink<T> sink;
  sink.ptr[kOutTx] = a.Tx;
  sink.ptr[kOutT] = a.Tm;
  sink.ptr[kOutR] = a.R;
  sink.ptr[kOutTinv] = a.Tinv;
  sink.ptr[kOutQuat] = a.quat;
  sink.ptr[kOutJ] = a.J;
  sink.ptr[kOutdJ] = a.dJ;
  sink.ptr[kOutM] = a.M;
  sink.ptr[kOutg] = a.g;
  sink.ptr[kOutC] = a.C;
  sink.tile = region + WS::kKin;
  sink.lane = lane;
  pdl_entry();
  for (int64_t base = (int64_t)blockIdx.x * kBlock; base < a.B; base += (int64_t)gridDim.x * kBlock) {
    // no early exit: every thread of the CTA takes part in the phase barriers; idle lanes / warps redo a valid state
    const int64_t warp_b0 = base + warp * 32;
    const int64_t rem = a.B - warp_b0;
    const int nvalid = rem < 32 ? (rem > 0 ? (int)rem : 0) : 32;
    const int64_t bb = warp_b0 + (lane < nvalid ? lane : nvalid - 1);
    const int64_t b = bb < a.B ? (bb >= 0 ? bb : 0) : a.B - 1;
    T q[N], dq[N];
#pragma unroll
    for (int k = 0; k < N; ++k) {
      q[k] = a.q[b * N + k];
      dq[k] = a.dq != nullptr ? a.dq[b * N + k] : T(0);
    }
    sink.warp_b0 = warp_b0;
    sink.nvalid = nvalid;
    rbd_state<T, N, DYN, CMAT, XTRA>(P, q, dq, a.frame, a.xoff, a.want, K, sink);
  }
}

template <typename T>
struct OscArgs {
  const T *q, *dq, *target, *tv;
  T *u, *train;
  T *ierr;  // (B, 6) integrated task-space error, updated in place (ki != 0), or nullptr
  int64_t B;
  int target_stride, tv_stride;
  GatherArgs g;  // n_peer > 0: also store u into every rank's gathered array (peer memory over NVLink)
  int *sched;    // {next tile, CTAs done} of this launch (zero on entry, re-armed by the last CTA), or nullptr: static tiles
};

// CTA-level shared memory of the OSC kernel behind the per-warp regions: the queue of deferred states (abrb_coop.cuh)
template <typename T, int N, int KD, int CAP>
struct OscQueue {
  static constexpr int kRecElems = CAP * CoopRecord<N, KD>::kLen;
  static constexpr size_t kRowOff = ((size_t)kRecElems * sizeof(T) + 15) / 16 * 16;   // long long rows[CAP]
  static constexpr size_t kCountOff = kRowOff + CAP * sizeof(long long);             // int count
  static constexpr size_t kBytes = kCountOff + 32;  // count (int), next tile (long long)
};

// One pass over the batch, persistent CTAs (grid = resident CTAs, tiles from a per-launch counter).  The states whose
// task-space inertia needs the truncating pseudo-inverse (3.8 % of uniformly random UR5 6-DOF states: 70 % of the warps
// hold one) leave a record in the CTA's queue and are finished by the whole CTA cooperatively (abrb_coop.cuh) once as
// many as the CTA has groups (20 or 16) have gathered or the CTA has run out of tiles; what does not fit the queue is
// finished by its warp in line.
// CTA size of the OSC kernel.  Measured on B200 (UR5 6-DOF fp64, B = 65 536): 128 threads 60.4 us, 64 threads 65.7 us —
// what ends the kernel is the CTA whose queue happens to hold more records than it has groups (a second Jacobi pass),
// and smaller CTAs have fewer groups per queue; 256 threads do not fit the non-orthonormal fp64 scratch.
#ifndef ABRB_OSC_BLOCK
#define ABRB_OSC_BLOCK 128
#endif
// The orthonormal-chain fp64 kernels (UR5 ...: 54 scratch slots per lane) could afford one 256-thread CTA per SM instead
// of two 128-thread ones (eight warps sharing the instruction fetches of the same straight-line code, a queue that
// practically never overflows the CTA's groups).  Measured on B200, UR5 6-DOF fp64 B = 65 536: 60.7 us against 60.4 us —
// the flush gets cheaper and the evaluation itself slower (40.1 against 37.0 us without pseudo-inverse states) — so both
// stay at 128; the non-orthonormal scratch (126 slots) does not fit eight warps anyway.
#ifndef ABRB_OSC_BLOCK_ORTHO64
#define ABRB_OSC_BLOCK_ORTHO64 128
#endif
// CTA phase barriers inside the evaluation (abrb_math.cuh, RegStore / StridedStore::sync)
#ifndef ABRB_OSC_PSYNC_F64
#define ABRB_OSC_PSYNC_F64 1
#endif
#ifndef ABRB_OSC_PSYNC_F32
#define ABRB_OSC_PSYNC_F32 0
#endif
template <typename T>
struct OscPhaseSync {
  static constexpr bool value = sizeof(T) == 8 ? (ABRB_OSC_PSYNC_F64 != 0) : (ABRB_OSC_PSYNC_F32 != 0);
};

template <typename T, bool ORTHO>
struct OscBlock {
  static constexpr int value = (ORTHO && sizeof(T) == 8) ? ABRB_OSC_BLOCK_ORTHO64 : ABRB_OSC_BLOCK;
};

// Resident CTAs per SM the OSC kernel is compiled for.  fp64: 2 (255 registers).  fp32 with orthonormal frames (UR5: the
// signed-permutation constants fold away): 4 (128 registers) measured best; fp32 with general frames (Jaco2) spills
// ~3 KB per thread at 128 registers — 2 CTAs at 255 registers run config 3 in 71.7 instead of 83.2 us and config 5 in
// 173 instead of 189 us (profiles/r02_experiments.md).
#ifndef ABRB_MINBLOCKS_OSC_F32_GENERAL
#define ABRB_MINBLOCKS_OSC_F32_GENERAL 2
#endif
template <typename T, bool ORTHO>
struct MinBlocksOsc {
  static constexpr int value = sizeof(T) == 8 ? MinBlocks<T>::value : (ORTHO ? MinBlocks<T>::value : ABRB_MINBLOCKS_OSC_F32_GENERAL);
};

template <typename T, int N, bool ORTHO, int KD, bool KSMEM>
__global__ void __launch_bounds__(OscBlock<T, ORTHO>::value,
                                  (MinBlocksOsc<T, ORTHO>::value * kBlock / OscBlock<T, ORTHO>::value > 0
                                       ? MinBlocksOsc<T, ORTHO>::value * kBlock / OscBlock<T, ORTHO>::value : 1))
osc_kernel(const __grid_constant__ ChainK<T, N> P, const __grid_constant__ OscK<T, N> O,
           const __grid_constant__ OscArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  typedef KinSel<T, N, ORTHO, KSMEM> KS;
  typedef OscSmem<T, N, ORTHO, KD, KSMEM> WS;
  constexpr int kOscBlock = OscBlock<T, ORTHO>::value, kOscWarps = kOscBlock / 32;
  constexpr int kOscFlushAt = kOscWarps * CoopGroup<N, KD>::kPerWarp;  // one full round of the CTA's groups
  constexpr int kCoopQueue = kCoopQueuePerWarp * kOscWarps;
  typedef OscQueue<T, N, KD, kCoopQueue> Q;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T *region = reinterpret_cast<T *>(smem_raw) + warp * WS::kElems;
  T *stage = region + WS::kKin;
  unsigned char *qbase = smem_raw + ((size_t)kOscWarps * WS::kElems * sizeof(T) + 15) / 16 * 16;
  T *qrec = reinterpret_cast<T *>(qbase);
  long long *qrow = reinterpret_cast<long long *>(qbase + Q::kRowOff);
  int *qcount = reinterpret_cast<int *>(qbase + Q::kCountOff);
  if (threadIdx.x == 0) *qcount = 0;
  __syncthreads();
#ifdef ABRB_DBG_TIMING  // (timing experiments only) per-CTA cycle counts are written over the training-signal buffer
  const long long dbg_t0 = clock64();
  long long dbg_flush = 0, dbg_wait = 0;
  int dbg_nflush = 0, dbg_rec = 0;
#endif
  typename KS::type K;
  KS::bind(K, region, lane);
  K.s.psync = OscPhaseSync<T>::value;
  WarpCoop<T, N, KD, typename KS::type> coop{region + WS::kKin + WS::kTile, region, lane, true, qrec, qrow, qcount, 0,
                                             kCoopQueue};
  FlushOut<T> fo;
  fo.u = a.u;
#ifdef ABRB_DBG_TIMING
  fo.train = nullptr;
#else
  fo.train = a.train;
#endif
  fo.n_peer = a.g.n_peer;
  fo.self = a.g.self;
  fo.row0 = a.g.row0;
#pragma unroll
  for (int p = 0; p < kMaxPeers; ++p) fo.peer[p] = static_cast<T *>(a.g.peer_u[p]);
  const double rcond = double(O.thr) * 0.1;
  // Programmatic dependent launch: everything above touched only kernel parameters and shared memory.  A launch that
  // follows another kernel in its stream is allowed onto the SMs while that kernel's last CTAs are still running (its
  // launch latency, parameter upload and CTA ramp-up overlap their tail) and waits HERE, before its first global access,
  // until that kernel has completed and its memory is visible.  Without the launch attribute both are no-ops.
  pdl_entry();
  // Tiles: the first one is the CTA's own index; the following ones come from the launch's tile counter, so that a CTA
  // whose tiles happen to be expensive (obstacle-active states, many pseudo-inverse states) simply takes fewer of them.
  const long long n_tiles = (a.B + kOscBlock - 1) / kOscBlock;
  long long *next_tile = reinterpret_cast<long long *>(qbase + Q::kCountOff + 8);
  for (long long tile = blockIdx.x; tile < n_tiles;) {
    const int64_t base = tile * kOscBlock;
    // (the next tile's index is asked for now and published at the end of this tile: the atomic's round trip hides
    // under the evaluation)
    long long upcoming = 0;
    if (threadIdx.x == 0) upcoming = a.sched != nullptr ? (long long)gridDim.x + atomicAdd(a.sched, 1) : tile + gridDim.x;
    // no early exit: every lane of a warp takes part in the cooperative steps; idle lanes / warps redo a valid state
    const int64_t warp_b0 = base + warp * 32;
    const int64_t rem = a.B - warp_b0;
    const int nvalid = rem < 32 ? (rem > 0 ? (int)rem : 0) : 32;
    const int64_t bb = warp_b0 + (lane < nvalid ? lane : nvalid - 1);
    const int64_t b = bb < a.B ? (bb >= 0 ? bb : 0) : a.B - 1;
    T q[N], dq[N], tg[6], tv[6], ie[6], u[N], tr[N];
#pragma unroll
    for (int k = 0; k < N; ++k) {
      q[k] = a.q[b * N + k];
      dq[k] = a.dq[b * N + k];
    }
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      tg[c] = a.target[b * a.target_stride + c];
      tv[c] = a.tv != nullptr ? a.tv[b * a.tv_stride + c] : T(0);
      ie[c] = a.ierr != nullptr ? a.ierr[b * 6 + c] : T(0);
    }
    coop.valid = lane < nvalid;
    coop.row = b;
    osc_eval<T, N, KD, false>(P, O, q, dq, tg, a.tv != nullptr ? tv : nullptr, a.ierr != nullptr ? ie : nullptr, u, tr,
                              (T *)nullptr, K, coop);
    if (a.u) store_records<T, N>(a.u, warp_b0, nvalid, u, stage, lane);
#ifndef ABRB_DBG_TIMING
    if (a.train) store_records<T, N>(a.train, warp_b0, nvalid, tr, stage, lane);
#endif
    if (a.ierr) store_records<T, 6>(a.ierr, warp_b0, nvalid, ie, stage, lane);
    // fused all-gather: this tile's rows go to every rank's gathered array while the other warps still compute
    for (int p = 0; p < a.g.n_peer; ++p)
      store_records<T, N>(static_cast<T *>(a.g.peer_u[p]), a.g.row0 + warp_b0, nvalid, u, stage, lane);
    // deferred states: emptied once a full round of the CTA's groups has gathered, and after the last tile
#ifdef ABRB_DBG_TIMING
    const long long dbg_w0 = clock64();
#endif
    if (threadIdx.x == 0) *next_tile = upcoming;
    __syncthreads();  // this tile's rows and records, and the next tile's index, are visible to the whole CTA
#ifdef ABRB_DBG_TIMING
    dbg_wait += clock64() - dbg_w0;
    const long long dbg_f0 = clock64();
#endif
    const int queued = *qcount < kCoopQueue ? *qcount : kCoopQueue;
    tile = *next_tile;
    const bool last = tile >= n_tiles;
    if (queued >= kOscFlushAt || (last && queued > 0)) {
#ifndef ABRB_DBG_NOFLUSH  // (timing experiments only: results of the deferred states are then wrong)
      coop_flush_cta<T, N, KD>(qrec, qrow, queued, fo, rcond, O.n_null > 0);
#endif
      __syncthreads();
      if (threadIdx.x == 0) *qcount = 0;
#ifdef ABRB_DBG_TIMING
      dbg_flush += clock64() - dbg_f0;
      dbg_nflush += 1;
      dbg_rec += queued;
#endif
    }
    __syncthreads();
  }
  if (a.sched != nullptr && threadIdx.x == 0) {  // every CTA has taken its last tile index: the last one re-arms the counter
    const int done = atomicAdd(a.sched + 1, 1);
    if (done == (int)gridDim.x - 1) {
      a.sched[0] = 0;
      a.sched[1] = 0;
      __threadfence();
    }
  }
#ifdef ABRB_DBG_TIMING
  if (lane == 0 && a.train != nullptr) {
    T *d = a.train + ((size_t)blockIdx.x * kOscWarps + warp) * 6;
    d[0] = T(clock64() - dbg_t0);
    d[1] = T(dbg_flush);
    d[2] = T(dbg_wait);
    d[3] = T(dbg_nflush);
    d[4] = T(dbg_rec);
    d[5] = T(gridDim.x);
  }
#endif
  if (a.g.n_peer > 0) {
    // completion: once every CTA's peer stores are visible system-wide, the last CTA publishes this launch's epoch
    // in every rank's flag array; abrb_gather_wait() on the consumer side spins on those flags
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned done = atomicAdd(a.g.cta_counter, 1u);
      if (done == gridDim.x - 1) {
        *a.g.cta_counter = 0u;
        __threadfence_system();
        for (int p = 0; p < a.g.n_peer; ++p)
          asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a.g.peer_flag[p]), "l"(a.g.epoch) : "memory");
      }
    }
  }
}

template <typename T>
struct RolloutArgs {
  T *q, *dq;
  const T *target;
  T *q_traj, *dq_traj, *u_traj;
  T *ierr;  // (B, 6) integrated task-space error, in/out (ki != 0), or nullptr
  int64_t B;
  int target_stride, steps;
  T dt;
};

// Closed loop: u = OSC(q, dq); ddq = M^-1 (u + g - C dq); dq += ddq dt; q += dq dt  (state stays in registers)
template <typename T, int N, bool ORTHO, int KD, bool KSMEM>
__global__ void __launch_bounds__(kBlock)
rollout_kernel(const __grid_constant__ ChainK<T, N> P, const __grid_constant__ OscK<T, N> O,
               const __grid_constant__ RolloutArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  typedef KinSel<T, N, ORTHO, KSMEM> KS;
  typedef OscSmem<T, N, ORTHO, KD, KSMEM> WS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T *region = reinterpret_cast<T *>(smem_raw) + warp * WS::kElems;
  T *stage = region + WS::kKin;
  typename KS::type K;
  KS::bind(K, region, lane);
  K.s.psync = OscPhaseSync<T>::value;
  WarpCoop<T, N, KD, typename KS::type> coop{region + WS::kKin + WS::kTile, region, lane, true};
  for (int64_t base = (int64_t)blockIdx.x * kBlock; base < a.B; base += (int64_t)gridDim.x * kBlock) {
    // no early exit (cooperative step inside osc_eval): idle lanes / warps redo a valid state and store nothing
    const int64_t warp_b0 = base + warp * 32;
    const int64_t rem = a.B - warp_b0;
    const int nvalid = rem < 32 ? (rem > 0 ? (int)rem : 0) : 32;
    const int64_t bb = warp_b0 + (lane < nvalid ? lane : nvalid - 1);
    const int64_t b = bb < a.B ? (bb >= 0 ? bb : 0) : a.B - 1;
    T q[N], dq[N], tg[6], ie[6], u[N], acc[N];
#pragma unroll
    for (int k = 0; k < N; ++k) {
      q[k] = a.q[b * N + k];
      dq[k] = a.dq[b * N + k];
    }
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      tg[c] = a.target[b * a.target_stride + c];
      ie[c] = a.ierr != nullptr ? a.ierr[b * 6 + c] : T(0);
    }
    coop.valid = lane < nvalid;
    for (int t = 0; t < a.steps; ++t) {
      osc_eval<T, N, KD, true>(P, O, q, dq, tg, (const T *)nullptr, a.ierr != nullptr ? ie : nullptr, u, (T *)nullptr,
                               acc, K, coop);
#pragma unroll
      for (int k = 0; k < N; ++k) {
        dq[k] += acc[k] * a.dt;
        q[k] += dq[k] * a.dt;
      }
      const int64_t row0 = (int64_t)t * a.B + warp_b0;
      if (a.u_traj) store_records<T, N>(a.u_traj, row0, nvalid, u, stage, lane);
      if (a.q_traj) store_records<T, N>(a.q_traj, row0, nvalid, q, stage, lane);
      if (a.dq_traj) store_records<T, N>(a.dq_traj, row0, nvalid, dq, stage, lane);
    }
    store_records<T, N>(a.q, warp_b0, nvalid, q, stage, lane);
    store_records<T, N>(a.dq, warp_b0, nvalid, dq, stage, lane);
    if (a.ierr) store_records<T, 6>(a.ierr, warp_b0, nvalid, ie, stage, lane);
  }
}

template <typename T>
struct NullArgs {
  const T *q, *dq;
  T *u;
  int64_t B;
};

template <typename T, int N, bool ORTHO>
__global__ void __launch_bounds__(kBlock)
null_kernel(const __grid_constant__ ChainK<T, N> P, const __grid_constant__ NullK<T, N> Z,
            const __grid_constant__ NullArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T *stage = reinterpret_cast<T *>(smem_raw) + warp * kPitch * (N < kChunk ? N : kChunk);
  for (int64_t base = (int64_t)blockIdx.x * kBlock; base < a.B; base += (int64_t)gridDim.x * kBlock) {
    const int64_t warp_b0 = base + warp * 32;
    if (warp_b0 >= a.B) break;
    const int64_t rem = a.B - warp_b0;
    const int nvalid = rem < 32 ? (int)rem : 32;
    const int64_t b = warp_b0 + (lane < nvalid ? lane : nvalid - 1);
    T q[N], dq[N], u[N];
#pragma unroll
    for (int k = 0; k < N; ++k) {
      q[k] = a.q[b * N + k];
      dq[k] = a.dq[b * N + k];
    }
    Kin<T, N, ORTHO> K;
    null_state<T, N>(P, Z, q, dq, u, K);
    store_records<T, N>(a.u, warp_b0, nvalid, u, stage, lane);
  }
}

template <typename T>
struct CtrlArgs {
  const T *q, *dq, *target, *tv;
  T *u;
  int64_t B;
  int target_stride, tv_stride, kind, flag_a, flag_b;
  T kp, kv;
};

// Joint.generate / Floating.generate: small relatives of the kernels above (M, g, one product or one 3x3 solve)
template <typename T, int N, bool ORTHO>
__global__ void __launch_bounds__(kBlock)
ctrl_kernel(const __grid_constant__ ChainK<T, N> P, const __grid_constant__ CtrlArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T *stage = reinterpret_cast<T *>(smem_raw) + warp * kPitch * N;
  for (int64_t base = (int64_t)blockIdx.x * kBlock; base < a.B; base += (int64_t)gridDim.x * kBlock) {
    const int64_t warp_b0 = base + warp * 32;
    if (warp_b0 >= a.B) break;
    const int64_t rem = a.B - warp_b0;
    const int nvalid = rem < 32 ? (int)rem : 32;
    const int64_t b = warp_b0 + (lane < nvalid ? lane : nvalid - 1);
    T q[N], dq[N], tg[N], tv[N], u[N];
#pragma unroll
    for (int k = 0; k < N; ++k) {
      q[k] = a.q[b * N + k];
      dq[k] = a.dq != nullptr ? a.dq[b * N + k] : T(0);
      tg[k] = a.target != nullptr ? a.target[b * a.target_stride + k] : T(0);
      tv[k] = a.tv != nullptr ? a.tv[b * a.tv_stride + k] : T(0);
    }
    Kin<T, N, ORTHO> K;
    if (a.kind == 0)
      joint_state<T, N>(P, a.kp, a.kv, a.flag_a != 0, q, dq, tg, a.tv != nullptr ? tv : nullptr, u, K);
    else
      floating_state<T, N>(P, a.flag_a != 0, a.flag_b != 0, q, dq, u, K);
    store_records<T, N>(a.u, warp_b0, nvalid, u, stage, lane);
  }
}

template <typename T>
struct SlidingArgs {
  const T *q, *dq, *target, *tv, *ta;
  T *u, *s;
  int64_t B;
  int target_stride, tv_stride, ta_stride, cartesian, frame;
  T kd, lamb;
  T xoff[3];
};

// Sliding.generate: J, dJ, M, C, g of one state and two applications of pinv(J) (abrb_osc.cuh, sliding_state)
template <typename T, int N, bool ORTHO>
__global__ void __launch_bounds__(kBlock)
sliding_kernel(const __grid_constant__ ChainK<T, N> P, const __grid_constant__ SlidingArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T *stage = reinterpret_cast<T *>(smem_raw) + warp * kPitch * N;
  const int w = a.cartesian ? 3 : N;
  for (int64_t base = (int64_t)blockIdx.x * kBlock; base < a.B; base += (int64_t)gridDim.x * kBlock) {
    const int64_t warp_b0 = base + warp * 32;
    if (warp_b0 >= a.B) break;
    const int64_t rem = a.B - warp_b0;
    const int nvalid = rem < 32 ? (int)rem : 32;
    const int64_t b = warp_b0 + (lane < nvalid ? lane : nvalid - 1);
    T q[N], dq[N], tg[N], tv[N], ta[N], u[N], sv[N];
#pragma unroll
    for (int k = 0; k < N; ++k) {
      q[k] = a.q[b * N + k];
      dq[k] = a.dq[b * N + k];
      const bool on = k < w;
      tg[k] = on ? a.target[b * a.target_stride + k] : T(0);
      tv[k] = (on && a.tv != nullptr) ? a.tv[b * a.tv_stride + k] : T(0);
      ta[k] = (on && a.ta != nullptr) ? a.ta[b * a.ta_stride + k] : T(0);
    }
    Kin<T, N, ORTHO> K;
    sliding_state<T, N>(P, a.kd, a.lamb, a.cartesian != 0, a.frame, a.xoff, q, dq, tg, tv, ta, u, sv, K);
    store_records<T, N>(a.u, warp_b0, nvalid, u, stage, lane);
    if (a.s != nullptr) store_records<T, N>(a.s, warp_b0, nvalid, sv, stage, lane);
  }
}

template <typename T>
struct IkArgs {
  const T *position, *target;
  T *pos_path, *vel_path;  // (steps, B, n)
  int64_t B;
  int target_stride, steps, method;
  T max_dx, max_dr, max_dq;  // already multiplied by dt
};

// InverseKinematics.generate_path: one trajectory per thread, the joint state stays in registers over the steps
// (sequential by construction, like the rollout kernel)
template <typename T, int N, bool ORTHO>
__global__ void __launch_bounds__(kBlock)
ik_kernel(const __grid_constant__ ChainK<T, N> P, const __grid_constant__ IkArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T *stage = reinterpret_cast<T *>(smem_raw) + warp * kPitch * N;
  for (int64_t base = (int64_t)blockIdx.x * kBlock; base < a.B; base += (int64_t)gridDim.x * kBlock) {
    const int64_t warp_b0 = base + warp * 32;
    if (warp_b0 >= a.B) break;
    const int64_t rem = a.B - warp_b0;
    const int nvalid = rem < 32 ? (int)rem : 32;
    const int64_t b = warp_b0 + (lane < nvalid ? lane : nvalid - 1);
    T q[N], dq[N], tg[3], Qd[4];
#pragma unroll
    for (int k = 0; k < N; ++k) q[k] = a.position[b * N + k];
#pragma unroll
    for (int c = 0; c < 3; ++c) tg[c] = a.target[b * a.target_stride + c];
    quat_from_euler_sxyz(a.target[b * a.target_stride + 3], a.target[b * a.target_stride + 4],
                         a.target[b * a.target_stride + 5], Qd);
    const T nq = T(1) / sqrt_t(Qd[0] * Qd[0] + Qd[1] * Qd[1] + Qd[2] * Qd[2] + Qd[3] * Qd[3]);
#pragma unroll
    for (int i = 0; i < 4; ++i) Qd[i] *= nq;
    for (int t = 0; t < a.steps; ++t) {
      Kin<T, N, ORTHO> K;
      ik_step<T, N>(P, a.max_dx, a.max_dr, a.max_dq, a.method, q, tg, Qd, dq, K);
      const int64_t row0 = (int64_t)t * a.B + warp_b0;
      store_records<T, N>(a.pos_path, row0, nvalid, q, stage, lane);
      store_records<T, N>(a.vel_path, row0, nvalid, dq, stage, lane);
#pragma unroll
      for (int k = 0; k < N; ++k) q[k] += dq[k];
    }
  }
}

// ------------------------------------------------------------------------------------------------ launch
// programmatic dependent launch of the OSC kernel (on unless ABRB_PDL=0 in the environment: an A/B switch for timing)
inline bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char *v = std::getenv("ABRB_PDL");
    on = (v != nullptr && v[0] == '0') ? 0 : 1;
  }
  return on == 1;
}

// kernel<<<grid, block, smem, stream>>>(args...) with programmatic stream serialization allowed: the launch may be
// brought onto the SMs while the stream's previous kernel drains; the kernels launched through it (osc_kernel,
// rbd_kernel) call pdl_entry() before their first global-memory access, which holds them until that kernel has
// completed and flushed.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t stream,
                              Args &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

inline int num_sms() {
  static int sm_count = 0;
  if (sm_count == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sm_count <= 0)
      sm_count = 148;
  }
  return sm_count;
}

inline int grid_for(int64_t B, int blocks_per_sm) {
  const int64_t tiles = (B + kBlock - 1) / kBlock;
  const int64_t cap = (int64_t)num_sms() * blocks_per_sm;
  return (int)(tiles < cap ? tiles : cap);
}

template <typename K>
inline cudaError_t set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  return cudaSuccess;
}

template <typename T, int N, bool ORTHO, bool DYN, bool CMAT, bool XTRA>
int rbd_go(const ChainHost &h, const RbdCall &c, unsigned want) {
  ChainK<T, N> P;
  fill_chain<T, N>(h, P);
  RbdArgs<T> a;
  a.q = static_cast<const T *>(c.q);
  a.dq = static_cast<const T *>(c.dq);
  a.Tx = static_cast<T *>(c.out.Tx);
  a.Tm = static_cast<T *>(c.out.T);
  a.R = static_cast<T *>(c.out.R);
  a.Tinv = static_cast<T *>(c.out.T_inv);
  a.quat = static_cast<T *>(c.out.quat);
  a.J = static_cast<T *>(c.out.J);
  a.dJ = static_cast<T *>(c.out.dJ);
  a.M = static_cast<T *>(c.out.M);
  a.g = static_cast<T *>(c.out.g);
  a.C = static_cast<T *>(c.out.C);
  a.B = c.B;
  a.frame = c.frame;
  a.want = want;
  for (int i = 0; i < 3; ++i) a.xoff[i] = c.xoff ? T(c.xoff[i]) : T(0);
  constexpr bool KSMEM = sizeof(T) == 8 || ABRB_KSMEM_F32 || ABRB_ROLLED;  // rolled loops index the scratch at run time
  const size_t smem = (size_t)kWarps * WarpSmem<T, N, ORTHO, KSMEM, MaxRecord<N>::value>::kElems * sizeof(T);
  auto kern = rbd_kernel<T, N, ORTHO, DYN, CMAT, XTRA, KSMEM>;
  cudaError_t e = set_smem(kern, smem);
  if (e != cudaSuccess) return (int)e;
  e = launch_pdl(kern, grid_for(c.B, 8), kBlock, smem, c.stream, P, a);
  count_launch();
  return e != cudaSuccess ? (int)e : (int)cudaGetLastError();
}

template <typename T, int N>
int rbd_dispatch(const ChainHost &h, const RbdCall &c) {
  unsigned want = 0;
  if (c.out.Tx) want |= kWantTx;
  if (c.out.T) want |= kWantT;
  if (c.out.R) want |= kWantR;
  if (c.out.T_inv) want |= kWantTinv;
  if (c.out.quat) want |= kWantQuat;
  if (c.out.J) want |= kWantJ;
  if (c.out.dJ) want |= kWantdJ | kWantJ;
  if (c.out.M) want |= kWantM;
  if (c.out.g) want |= kWantg;
  if (c.out.C) want |= kWantC;
  const bool cm = c.out.C != nullptr, dyn = c.out.M != nullptr || c.out.g != nullptr;
  const bool xtra = (want & ~(kWantJ | kWantM | kWantg | kWantC)) != 0;
  // instantiations: {frame-only (with extras)} + {dynamics / dynamics+C} x {with, without extras}
#define ABRB_RBD_GO(O_)                                                                   \
  do {                                                                                     \
    if (cm) return xtra ? rbd_go<T, N, O_, true, true, true>(h, c, want) : rbd_go<T, N, O_, true, true, false>(h, c, want);   \
    if (dyn) return xtra ? rbd_go<T, N, O_, true, false, true>(h, c, want) : rbd_go<T, N, O_, true, false, false>(h, c, want); \
    return rbd_go<T, N, O_, false, false, true>(h, c, want);                               \
  } while (0)
  if (h.ortho) ABRB_RBD_GO(true);
  ABRB_RBD_GO(false);
#undef ABRB_RBD_GO
}

template <typename T, int N, bool ORTHO, int KD>
int osc_go(const ChainHost &h, const abrb_osc_params &p, const OscCall &c) {
  ChainK<T, N> P;
  fill_chain<T, N>(h, P);
  OscK<T, N> O;
  fill_osc<T, N>(p, c.frame, c.xoff, O);
  OscArgs<T> a;
  a.q = static_cast<const T *>(c.q);
  a.dq = static_cast<const T *>(c.dq);
  a.target = static_cast<const T *>(c.target);
  a.tv = static_cast<const T *>(c.tv);
  a.u = static_cast<T *>(c.u);
  a.train = static_cast<T *>(c.train);
  a.ierr = static_cast<T *>(c.ierr);
  a.B = c.B;
  a.target_stride = c.target_stride;
  a.tv_stride = c.tv_stride;
  if (c.gather != nullptr) a.g = *c.gather;
  a.sched = c.sched;
  constexpr bool KSMEM = sizeof(T) == 8 || ABRB_KSMEM_F32 || ABRB_ROLLED;  // rolled loops index the scratch at run time
  constexpr int kOscBlock = OscBlock<T, ORTHO>::value, kOscWarps = kOscBlock / 32;
  const size_t smem = ((size_t)kOscWarps * OscSmem<T, N, ORTHO, KD, KSMEM>::kElems * sizeof(T) + 15) / 16 * 16 +
                      OscQueue<T, N, KD, kCoopQueuePerWarp * kOscWarps>::kBytes;
  auto kern = osc_kernel<T, N, ORTHO, KD, KSMEM>;
  cudaError_t e = set_smem(kern, smem);
  if (e != cudaSuccess) return (int)e;
  // persistent CTAs: as many as are resident at once, so that each sees several tiles and its deferred states gather
  static thread_local int resident = 0;
  static thread_local int resident_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (resident_dev != dev) {
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kOscBlock, smem);
    if (e != cudaSuccess) return (int)e;
    resident = per_sm > 0 ? per_sm : 1;
    resident_dev = dev;
  }
  const int64_t tiles = (c.B + kOscBlock - 1) / kOscBlock, cap = (int64_t)num_sms() * resident;
  e = launch_pdl(kern, (unsigned)(tiles < cap ? tiles : cap), kOscBlock, smem, c.stream, P, O, a);
  count_launch();
  return e != cudaSuccess ? (int)e : (int)cudaGetLastError();
}

template <typename T, int N, bool ORTHO, int KD>
int rollout_go(const ChainHost &h, const abrb_osc_params &p, const RolloutCall &c) {
  ChainK<T, N> P;
  fill_chain<T, N>(h, P);
  OscK<T, N> O;
  fill_osc<T, N>(p, c.frame, c.xoff, O);
  RolloutArgs<T> a;
  a.q = static_cast<T *>(c.q);
  a.dq = static_cast<T *>(c.dq);
  a.target = static_cast<const T *>(c.target);
  a.q_traj = static_cast<T *>(c.q_traj);
  a.dq_traj = static_cast<T *>(c.dq_traj);
  a.u_traj = static_cast<T *>(c.u_traj);
  a.ierr = static_cast<T *>(c.ierr);
  a.B = c.B;
  a.target_stride = c.target_stride;
  a.steps = c.steps;
  a.dt = T(c.dt);
  constexpr bool KSMEM = sizeof(T) == 8 || ABRB_KSMEM_F32 || ABRB_ROLLED;  // rolled loops index the scratch at run time
  const size_t smem = (size_t)kWarps * OscSmem<T, N, ORTHO, KD, KSMEM>::kElems * sizeof(T);
  auto kern = rollout_kernel<T, N, ORTHO, KD, KSMEM>;
  cudaError_t e = set_smem(kern, smem);
  if (e != cudaSuccess) return (int)e;
  kern<<<grid_for(c.B, 8), kBlock, smem, c.stream>>>(P, O, a);
  count_launch();
  return (int)cudaGetLastError();
}

template <typename T, int N, bool ORTHO>
int null_go(const ChainHost &h, const abrb_null_params &z, const NullCall &c) {
  ChainK<T, N> P;
  fill_chain<T, N>(h, P);
  NullK<T, N> Z;
  fill_null<T, N>(z, Z);
  NullArgs<T> a;
  a.q = static_cast<const T *>(c.q);
  a.dq = static_cast<const T *>(c.dq);
  a.u = static_cast<T *>(c.u);
  a.B = c.B;
  const size_t smem = (size_t)kWarps * kPitch * (N < kChunk ? N : kChunk) * sizeof(T);
  null_kernel<T, N, ORTHO><<<grid_for(c.B, 8), kBlock, smem, c.stream>>>(P, Z, a);
  count_launch();
  return (int)cudaGetLastError();
}

inline bool needs_kd6(const abrb_osc_params &p) { return p.ctrlr_dof[3] || p.ctrlr_dof[4] || p.ctrlr_dof[5]; }

}  // namespace

template <typename T, int N, bool ORTHO>
int ctrl_go(const ChainHost &h, const CtrlCall &c) {
  ChainK<T, N> P;
  fill_chain<T, N>(h, P);
  CtrlArgs<T> a;
  a.q = static_cast<const T *>(c.q);
  a.dq = static_cast<const T *>(c.dq);
  a.target = static_cast<const T *>(c.target);
  a.tv = static_cast<const T *>(c.tv);
  a.u = static_cast<T *>(c.u);
  a.B = c.B;
  a.target_stride = c.target_stride;
  a.tv_stride = c.tv_stride;
  a.kind = c.kind;
  a.flag_a = c.flag_a;
  a.flag_b = c.flag_b;
  a.kp = T(c.kp);
  a.kv = T(c.kv);
  const size_t smem = (size_t)kWarps * kPitch * N * sizeof(T);
  ctrl_kernel<T, N, ORTHO><<<grid_for(c.B, 8), kBlock, smem, c.stream>>>(P, a);
  count_launch();
  return (int)cudaGetLastError();
}

template <>
int launch_rbd<ABRB_N>(const ChainHost &h, const RbdCall &c) {
  return c.f32 ? rbd_dispatch<float, ABRB_N>(h, c) : rbd_dispatch<double, ABRB_N>(h, c);
}

#define ABRB_OSC_DISPATCH(GO, ...)                                                              \
  do {                                                                                          \
    const bool k6 = needs_kd6(p);                                                               \
    if (c.f32) {                                                                                \
      if (h.ortho) return k6 ? GO<float, ABRB_N, true, 6>(__VA_ARGS__) : GO<float, ABRB_N, true, 3>(__VA_ARGS__);    \
      return k6 ? GO<float, ABRB_N, false, 6>(__VA_ARGS__) : GO<float, ABRB_N, false, 3>(__VA_ARGS__);               \
    }                                                                                           \
    if (h.ortho) return k6 ? GO<double, ABRB_N, true, 6>(__VA_ARGS__) : GO<double, ABRB_N, true, 3>(__VA_ARGS__);    \
    return k6 ? GO<double, ABRB_N, false, 6>(__VA_ARGS__) : GO<double, ABRB_N, false, 3>(__VA_ARGS__);               \
  } while (0)

template <>
int launch_osc<ABRB_N>(const ChainHost &h, const abrb_osc_params &p, const OscCall &c) {
  ABRB_OSC_DISPATCH(osc_go, h, p, c);
}

template <>
int launch_rollout<ABRB_N>(const ChainHost &h, const abrb_osc_params &p, const RolloutCall &c) {
  ABRB_OSC_DISPATCH(rollout_go, h, p, c);
}

template <>
int launch_null<ABRB_N>(const ChainHost &h, const abrb_null_params &z, const NullCall &c) {
  if (c.f32) return h.ortho ? null_go<float, ABRB_N, true>(h, z, c) : null_go<float, ABRB_N, false>(h, z, c);
  return h.ortho ? null_go<double, ABRB_N, true>(h, z, c) : null_go<double, ABRB_N, false>(h, z, c);
}

template <typename T, int N, bool ORTHO>
int sliding_go(const ChainHost &h, const SlidingCall &c) {
  ChainK<T, N> P;
  fill_chain<T, N>(h, P);
  SlidingArgs<T> a;
  a.q = static_cast<const T *>(c.q);
  a.dq = static_cast<const T *>(c.dq);
  a.target = static_cast<const T *>(c.target);
  a.tv = static_cast<const T *>(c.tv);
  a.ta = static_cast<const T *>(c.ta);
  a.u = static_cast<T *>(c.u);
  a.s = static_cast<T *>(c.s);
  a.B = c.B;
  a.target_stride = c.target_stride;
  a.tv_stride = c.tv_stride;
  a.ta_stride = c.ta_stride;
  a.cartesian = c.cartesian;
  a.frame = c.frame;
  a.kd = T(c.kd);
  a.lamb = T(c.lamb);
  for (int i = 0; i < 3; ++i) a.xoff[i] = c.xoff ? T(c.xoff[i]) : T(0);
  const size_t smem = (size_t)kWarps * kPitch * N * sizeof(T);
  sliding_kernel<T, N, ORTHO><<<grid_for(c.B, 8), kBlock, smem, c.stream>>>(P, a);
  count_launch();
  return (int)cudaGetLastError();
}

template <typename T, int N, bool ORTHO>
int ik_go(const ChainHost &h, const IkCall &c) {
  ChainK<T, N> P;
  fill_chain<T, N>(h, P);
  IkArgs<T> a;
  a.position = static_cast<const T *>(c.position);
  a.target = static_cast<const T *>(c.target);
  a.pos_path = static_cast<T *>(c.pos_path);
  a.vel_path = static_cast<T *>(c.vel_path);
  a.B = c.B;
  a.target_stride = c.target_stride;
  a.steps = c.steps;
  a.method = c.method;
  a.max_dx = T(c.max_dx * c.dt);
  a.max_dr = T(c.max_dr * c.dt);
  a.max_dq = T(c.max_dq * c.dt);
  const size_t smem = (size_t)kWarps * kPitch * N * sizeof(T);
  ik_kernel<T, N, ORTHO><<<grid_for(c.B, 8), kBlock, smem, c.stream>>>(P, a);
  count_launch();
  return (int)cudaGetLastError();
}

template <>
int launch_ik<ABRB_N>(const ChainHost &h, const IkCall &c) {
  if (c.f32) return h.ortho ? ik_go<float, ABRB_N, true>(h, c) : ik_go<float, ABRB_N, false>(h, c);
  return h.ortho ? ik_go<double, ABRB_N, true>(h, c) : ik_go<double, ABRB_N, false>(h, c);
}

template <>
int launch_sliding<ABRB_N>(const ChainHost &h, const SlidingCall &c) {
  if (c.f32) return h.ortho ? sliding_go<float, ABRB_N, true>(h, c) : sliding_go<float, ABRB_N, false>(h, c);
  return h.ortho ? sliding_go<double, ABRB_N, true>(h, c) : sliding_go<double, ABRB_N, false>(h, c);
}

template <>
int launch_ctrl<ABRB_N>(const ChainHost &h, const CtrlCall &c) {
  if (c.f32) return h.ortho ? ctrl_go<float, ABRB_N, true>(h, c) : ctrl_go<float, ABRB_N, false>(h, c);
  return h.ortho ? ctrl_go<double, ABRB_N, true>(h, c) : ctrl_go<double, ABRB_N, false>(h, c);
}

}  // namespace abrb
