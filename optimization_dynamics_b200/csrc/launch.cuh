// Launch heuristics + per-model launch entry points of the contact step / rollout kernels.  Every model's kernels are instantiated
// in ONE translation unit of their own (csrc/inst/contact_<model>.cu) so that liboptdyn_b200.so builds in parallel (the generated
// model code × lane configurations is what nvcc spends its minutes on); optdyn_b200.cu only sees the declarations below.
#pragma once
#include <cuda_runtime.h>
#include <stdlib.h>
#include "models.cuh"

namespace od {

template <class M, int G, int PPB, bool REG, bool BSYNC = false>
static inline cudaError_t launch_contact_cfg(const StepArgs& a, cudaStream_t s) {
    const int grid = (a.B + PPB - 1) / PPB;
    constexpr size_t smem = sizeof(double) * PPB * ContactIP<M, G, PPB, REG>::WS;
    if (smem > 48 * 1024) {   // > 48 KB of dynamic shared memory needs an explicit opt-in; per device, so set at every launch
        cudaError_t e = cudaFuncSetAttribute(contact_step_kernel<M, G, PPB, REG, BSYNC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    contact_step_kernel<M, G, PPB, REG, BSYNC><<<grid, G * PPB, smem, s>>>(a);
    return cudaGetLastError();
}
// OD_BSYNC (default 0 = off): block-phased execution for the models whose loop does not fit the instruction caches (planar push).
// Measured on B200 (profiles/r02g_*, 25 600 problems): the barrier-aligned warps do share their instruction fetches (stall cycles per
// issue waiting for instructions 5.97 → 0.15), but a 256-thread block lasts as long as the slowest of its 32 problems and holds its
// SM while the finished warps wait at the barrier (5.2 barrier-stall cycles per issue): 6.53 → 6.62 ms, no gain.  OD_BSYNC=1 turns it
// on from 4097 problems, OD_BSYNC=n from n problems.
inline int bsync_min_batch() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("OD_BSYNC"); v = e ? atoi(e) : 0; if (v == 1) v = 4097; }
    return v;
}

// Lanes per problem: 1 = one thread per problem (throughput configuration, large batches; LU in shared memory); 4 / 8 =
// cooperative groups (latency configuration: a 4096-problem batch alone would put a single warp on each SM) with the
// register-resident Gauss–Jordan of group_gj.cuh.  OD_LANES overrides the heuristic; OD_REG=0 forces the shared-memory LU.
inline int lanes_for(int B, bool heavy = false) {
    static int forced = -1;
    if (forced < 0) { const char* e = getenv("OD_LANES"); forced = e ? atoi(e) : 0; }
    if (forced == 1 || forced == 2 || forced == 4 || forced == 8 || forced == 16) return forced;
    // planar push (20×20 reduced system, 35 variables): measured 1024 problems 5.2 / 4.1 / 3.4 ms with 4 / 8 / 16 lanes,
    // 25 600 problems 12.0 / 10.0 ms with 4 / 8 lanes (register path; shared-memory LU: 5.3 and 12.0 ms)
    if (heavy) return B <= 4096 ? 16 : 8;
    // measured on B200 (hopper, r02t): 16 lanes win up to ≈ 1536 problems (every warp alone on its scheduler: the shorter per-lane chain
    // is pure latency: 1 problem 0.0270 / 0.0332 / 0.0352 ms with 16 / 8 / 4 lanes, 512: 0.0414 / 0.0476 / 0.0536, 1536: 0.0540 / 0.0577 /
    // 0.0567), 4 lanes from 2048 problems on (2048: 0.0557 vs 0.0576 with 16; 4096: 0.0624 / 0.0709 / 0.0987 with 4 / 8 / 16);
    // 8 lanes are never the best choice with the solution gather through the mirror
    if (B <= 1536) return 16;
    return 4;                      // 262144 problems: 113 M solves/s with 4 lanes (register path)
}
#ifndef OD_BUILD_PHASED_4LANE
#define OD_BUILD_PHASED_4LANE 0
#endif
inline int phased_min_batch() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("OD_PHASED"); v = e ? atoi(e) : 0; }
    return v;
}
inline bool reg_path() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("OD_REG"); v = e ? atoi(e) : 1; }
    return v != 0;
}

// Kernels launched by the last launch_contact call of this thread (the persistent sweep is two to four kernels: sweep, resume, IFT of the finished, IFT of the parked): od_launch_count bookkeeping.
inline int& last_launch_kernels() { static thread_local int n = 1; return n; }

// OD_PERSIST (default 1): persistent block-phased sweep + separate IFT kernel for the models with the rank-revealing IFT, from
// this many problems on (1 = default threshold 3072, 0 = never, n = from n problems; measured, profiles/r02za_*: 2048 problems 1.26 ms
// per-warp kernel / 1.31 ms sweep, 3072: 1.44 / 1.36, 4096: 1.52 / 1.33).  Needs the scratch the C ABI layer provides (work_queue,
// z_snapshots) and no fused gather.
inline int persist_min_batch() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("OD_PERSIST"); v = e ? atoi(e) : 1; if (v == 1) v = 3072; }
    return v;
}
// OD_PARK_ITER (default 16; 0 = off): problems of the persistent sweep that are unfinished after this many iterations are parked and
// continued by a second launch — all of them at once, one warp per block, 16 lanes per problem (StepArgs::park_iter).
inline int park_iter_default() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("OD_PARK_ITER"); v = e ? atoi(e) : 16; if (v < 0) v = 0; }
    return v;
}
// OD_TAIL_OVERLAP (default 1): resume of the parked problems and IFT of the finished ones side by side on two streams; 0 = one after the other
inline bool tail_overlap() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("OD_TAIL_OVERLAP"); v = e ? atoi(e) : 1; }
    return v != 0;
}
template <class M, int G, int PPB>
static inline cudaError_t launch_contact_persistent(const StepArgs& a0, cudaStream_t s) {
    StepArgs a = a0;
    a.resume = 0;
    a.park_iter = (a.z_park && a.park_list && a.park_info && park_iter_default() < a.opts.max_iter) ? park_iter_default() : 0;
    typedef ContactIP<M, G, PPB, true> IP;
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
    constexpr size_t smem = sizeof(double) * PPB * IP::WS;
    static_assert(smem <= 227 * 1024, "sweep kernel: shared memory per block");
    cudaError_t e = cudaMemsetAsync(a.work_queue, 0, 4 * sizeof(unsigned int), s);   // [sweep queue, parked count, resume queue, -]
    if (e != cudaSuccess) return e;
    if (smem > 48 * 1024) {
        if ((e = cudaFuncSetAttribute(contact_sweep_kernel<M, G, PPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    }
    static int per_sm = 0;                                    // (per instantiation; one process drives one GPU)
    if (!per_sm) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, contact_sweep_kernel<M, G, PPB>, G * PPB, smem);
        if (per_sm < 1) per_sm = 1;
    }
    int grid = sms * per_sm;
    const int need = (a.B + PPB - 1) / PPB;
    if (grid > need) grid = need;
    const bool grad = a.want_grad && a.dq1;
    if (a.park_iter > 0 && grad) {                            // progress words double as the "parked" marks of the IFT role below
        if ((e = cudaMemsetAsync(a.park_info, 0, 2 * sizeof(int) * (size_t)a.B, s)) != cudaSuccess) return e;
    }
    contact_sweep_kernel<M, G, PPB><<<grid, G * PPB, smem, s>>>(a);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    last_launch_kernels() = 1;
    constexpr int PPB2 = 32 / G;                              // IFT: one warp per block — short, and phased by construction
    constexpr size_t smem2 = sizeof(double) * PPB2 * ContactIP<M, G, PPB2, true>::WS;
    const int ift_blocks = (a.B + PPB2 - 1) / PPB2;
    if (a.park_iter > 0) {
        // resume: a whole warp per parked problem (32 lanes: one matrix row per lane; loop 5866 instructions against 6648 with 16
        // lanes), up to 8 one-warp blocks per SM; blocks without work leave at once — the parked count is only known on the device
        constexpr int RG = 32, RPPB = 1;
        constexpr size_t smem_r = sizeof(double) * RPPB * ContactIP<M, RG, RPPB, true>::WS;
        StepArgs b = a;
        b.park_iter = 0; b.resume = 1;
        int grid_r = sms * 8;
        if (grid_r > a.B) grid_r = a.B;
        if (smem_r > 48 * 1024) {
            if ((e = cudaFuncSetAttribute(contact_sweep_kernel<M, RG, RPPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_r)) != cudaSuccess) return e;
        }
        if (grad && smem2 > 48 * 1024) {
            if ((e = cudaFuncSetAttribute(contact_ift_kernel<M, G, PPB2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2)) != cudaSuccess) return e;
            if ((e = cudaFuncSetAttribute(contact_ift_kernel<M, G, PPB2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2)) != cudaSuccess) return e;
            if ((e = cudaFuncSetAttribute(contact_ift_kernel<M, G, PPB2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2)) != cudaSuccess) return e;
        }
        if (grad && a.side_stream && a.ev_fork && a.ev_join && tail_overlap()) {
            // fork: the resume launch stays on the caller's stream, the IFT of everything that finished in the sweep runs beside it on
            // the handle's second stream (parked problems are walked through dry); join; then the IFT of the parked problems
            cudaStream_t side = (cudaStream_t)a.side_stream;
            if ((e = cudaEventRecord((cudaEvent_t)a.ev_fork, s)) != cudaSuccess) return e;
            if ((e = cudaStreamWaitEvent(side, (cudaEvent_t)a.ev_fork, 0)) != cudaSuccess) return e;
            contact_sweep_kernel<M, RG, RPPB><<<grid_r, RG * RPPB, smem_r, s>>>(b);
            if ((e = cudaGetLastError()) != cudaSuccess) return e;
            contact_ift_kernel<M, G, PPB2, 1><<<ift_blocks, G * PPB2, smem2, side>>>(b);
            if ((e = cudaGetLastError()) != cudaSuccess) return e;
            if ((e = cudaEventRecord((cudaEvent_t)a.ev_join, side)) != cudaSuccess) return e;
            if ((e = cudaStreamWaitEvent(s, (cudaEvent_t)a.ev_join, 0)) != cudaSuccess) return e;
            contact_ift_kernel<M, G, PPB2, 2><<<(sms < ift_blocks ? sms : ift_blocks), G * PPB2, smem2, s>>>(b);
            if ((e = cudaGetLastError()) != cudaSuccess) return e;
            last_launch_kernels() = 4;
        } else {
            contact_sweep_kernel<M, RG, RPPB><<<grid_r, RG * RPPB, smem_r, s>>>(b);
            if ((e = cudaGetLastError()) != cudaSuccess) return e;
            last_launch_kernels() = 2;
            if (grad) {                                       // no second stream (or OD_TAIL_OVERLAP=0): resume, then the IFT of all problems
                contact_ift_kernel<M, G, PPB2, 0><<<ift_blocks, G * PPB2, smem2, s>>>(a);
                if ((e = cudaGetLastError()) != cudaSuccess) return e;
                last_launch_kernels() = 3;
            }
        }
    } else if (grad) {
        if (smem2 > 48 * 1024) {
            if ((e = cudaFuncSetAttribute(contact_ift_kernel<M, G, PPB2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2)) != cudaSuccess) return e;
        }
        contact_ift_kernel<M, G, PPB2, 0><<<ift_blocks, G * PPB2, smem2, s>>>(a);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        last_launch_kernels() = 2;
    }
    return cudaSuccess;
}

// WIDE: models large enough for 8 lanes; REGOK: models whose IFT runs on the register path (not the rank-revealing one)
template <class M, bool WIDE, bool REGOK>
static inline cudaError_t launch_contact(const StepArgs& a, cudaStream_t s) {
    const int lanes = lanes_for(a.B, M::ROBUST_IFT);
    last_launch_kernels() = 1;
    if constexpr (REGOK) {
        if (reg_path()) {
            if constexpr (WIDE && M::ROBUST_IFT) {
                if (persist_min_batch() > 0 && a.B >= persist_min_batch() && a.n_peers <= 1 && a.work_queue && (!a.want_grad || !a.dq1 || a.z_snapshots)) {
                    // 8 lanes at every size (measured, r02z: 4096 problems 1.37 ms against 1.64 with 16 lanes, 25 600: 3.05 / 4.05)
                    return launch_contact_persistent<M, 8, 32>(a, s);
                }
                // a gather with peers keeps one-warp blocks (the fused barrier counts blocks as they finish; rows should leave early)
                if (bsync_min_batch() > 0 && a.B >= bsync_min_batch() && lanes == 8 && a.n_peers <= 1) return launch_contact_cfg<M, 8, 32, true, true>(a, s);
            }
#if OD_BUILD_PHASED_4LANE
            // A/B (rejected, profiles/r02ze_*): the 4-lane hopper kernel in block-phased 128-thread blocks — shared instruction fetches, but
            // the votes become barriers: 4096 problems 0.0586 -> 0.0771 ms, 262 144: 2.28 -> 2.47 ms.  -DOD_BUILD_PHASED_4LANE=1 + OD_PHASED=n.
            if constexpr (WIDE && !M::ROBUST_IFT) {
                if (lanes == 4 && phased_min_batch() > 0 && a.B >= phased_min_batch() && a.n_peers <= 1) return launch_contact_cfg<M, 4, 32, true, true>(a, s);
            }
#endif
            if constexpr (WIDE) { if (lanes == 16) return launch_contact_cfg<M, 16, 2, true>(a, s); }
            if constexpr (WIDE) { if (lanes == 8) return launch_contact_cfg<M, 8, 4, true>(a, s); }
            if (lanes >= 4) return launch_contact_cfg<M, 4, 8, true>(a, s);
        }
    }
    if constexpr (WIDE) { if (lanes == 8) return launch_contact_cfg<M, 8, 4, false>(a, s); }
    if (lanes >= 4) return launch_contact_cfg<M, 4, 8, false>(a, s);
    if constexpr (WIDE) { if (lanes == 2) return launch_contact_cfg<M, 2, 16, false>(a, s); }
    return launch_contact_cfg<M, 1, 32, false>(a, s);
}

template <class M, int G, int PPB, bool REG>
static inline cudaError_t launch_rollout_cfg(const RolloutArgs& a, cudaStream_t s) {
    const int grid = (a.R + PPB - 1) / PPB;
    constexpr size_t smem = sizeof(double) * PPB * ContactIP<M, G, PPB, REG>::WS;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(contact_rollout_kernel<M, G, PPB, REG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    contact_rollout_kernel<M, G, PPB, REG><<<grid, G * PPB, smem, s>>>(a);
    return cudaGetLastError();
}
// rollouts are latency-bound for any realistic count (≤ a few thousand): cooperative lanes, register path where the model has it
template <class M, bool WIDE, bool REGOK>
static inline cudaError_t launch_rollout(const RolloutArgs& a, cudaStream_t s) {
    const int lanes = lanes_for(a.R, M::ROBUST_IFT);
    if constexpr (REGOK) {
        if (reg_path()) {
            if constexpr (WIDE) { if (lanes == 16) return launch_rollout_cfg<M, 16, 2, true>(a, s); }
            if constexpr (WIDE) { if (lanes == 8) return launch_rollout_cfg<M, 8, 4, true>(a, s); }
            return launch_rollout_cfg<M, 4, 8, true>(a, s);
        }
    }
    return launch_rollout_cfg<M, 4, 8, false>(a, s);
}


// per-model entry points (defined by OD_INSTANTIATE_CONTACT in csrc/inst/contact_<model>.cu)
#define OD_DECLARE_CONTACT(NAME)                                                     \
    cudaError_t od_launch_step_##NAME(const StepArgs& a, cudaStream_t s);            \
    cudaError_t od_launch_rollout_##NAME(const RolloutArgs& a, cudaStream_t s);
OD_DECLARE_CONTACT(acrobot_impact)
OD_DECLARE_CONTACT(acrobot_nominal)
OD_DECLARE_CONTACT(cartpole_friction)
OD_DECLARE_CONTACT(cartpole_frictionless)
OD_DECLARE_CONTACT(planar_push)
OD_DECLARE_CONTACT(hopper)
#undef OD_DECLARE_CONTACT

#define OD_INSTANTIATE_STEP(NAME, MODEL, WIDE, REGOK) \
    cudaError_t od_launch_step_##NAME(const StepArgs& a, cudaStream_t s) { return launch_contact<MODEL, WIDE, REGOK>(a, s); }
#define OD_INSTANTIATE_ROLLOUT(NAME, MODEL, WIDE, REGOK) \
    cudaError_t od_launch_rollout_##NAME(const RolloutArgs& a, cudaStream_t s) { return launch_rollout<MODEL, WIDE, REGOK>(a, s); }
#define OD_INSTANTIATE_CONTACT(NAME, MODEL, WIDE, REGOK) OD_INSTANTIATE_STEP(NAME, MODEL, WIDE, REGOK) OD_INSTANTIATE_ROLLOUT(NAME, MODEL, WIDE, REGOK)

}  // namespace od
