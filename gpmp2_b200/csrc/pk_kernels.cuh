// pk_kernels.cuh -- the PHASE-KERNEL PIPELINE of the Levenberg-Marquardt optimizer (vector-state robots, default
// factor set): the fused one-kernel optimizer of optimizer_kernel.cuh cut at its phase boundaries into three kernels,
//
//   pk_lin_kernel    configuration-parallel half of NonlinearFactorGraph::linearize: forward kinematics, SDF lookups,
//                    hinge, whitened per-configuration J^T J / J^T e  ->  global M-list           (registers: as many
//                    as a lane needs for a 7-DOF chain; NO shared-memory H, so occupancy is set by registers alone)
//   pk_solve_kernel  entry-parallel half (constant template + per-state pass + obstacle blocks from the M-list), then
//                    the damped two-sided block-tridiagonal Cholesky solve  ->  delta, linearized cost change
//                    (shared memory: one H; registers: half of the fused kernel's)
//   pk_err_kernel    NonlinearFactorGraph::error at x + delta, LevenbergMarquardtOptimizer::tryLambda's accept /
//                    reject, gpmp2::optimize's checkConvergence, outputs                  (no H, few registers: the
//                    SDF gather latency is hidden by 20+ resident warps instead of one warp's own pipelining)
//
// with the per-trajectory state (states, delta, lambda, errors, counters) in global memory between them.  One ROUND =
// one tryLambda of every active trajectory: lin (only the trajectories whose last step was accepted) -> solve -> err.
// Work lists carry the active trajectories from round to round; the host launches the fixed maximum number of rounds
// (2 max_iter + 3: every rejected step multiplies lambda by 10 below its bound 1e5, every accepted one divides it), the
// kernels of a round without work exit at once.
//
// Why (ncu of the fused kernel, profiles/r1z_*): 255 registers and 32 KB of shared memory per one-warp block hold the
// SM at 7 warps (issue slots 32 % busy), and 142 KB of SASS with 7 warps in 7 different phases miss the 32 KB L1.5
// instruction cache 12.6 % of the time -- the per-GPC instruction path (gcc) was 87 % busy, which is why an eighth warp
// bought nothing.  Each phase kernel is small enough for the instruction cache, and its occupancy is set by what THAT
// phase needs.  The arithmetic (and its association order) is the fused kernel's: results are bit-identical.
//
// Replaces the same reference code as optimizer_kernel.cuh: internal::BatchTrajOptimize
// (gpmp2/planner/BatchTrajOptimizer-inl.h:19-84), gpmp2::optimize (BatchTrajOptimizer.cpp:212-308) and GTSAM's
// LevenbergMarquardtOptimizer underneath.
#pragma once
#include "optimizer_kernel.cuh"

// 1: the solve kernel keeps only the diagonal blocks of H in shared memory and streams the coupling blocks
// (VecOpt::solve_streamed); 0: the whole H in shared memory as in the fused kernel
#ifndef PK_STREAMED_SOLVE
#define PK_STREAMED_SOLVE 1
#endif

namespace pk {

__device__ __forceinline__ const int32_t* list_of(const KProblem& pr, int parity, int which) {
  return pr.pk_lists + (size_t)(parity * 2 + which) * pr.B;
}
__device__ __forceinline__ int32_t* list_of_mut(const KProblem& pr, int parity, int which) {
  return pr.pk_lists + (size_t)(parity * 2 + which) * pr.B;
}

// persistent one-warp blocks pull list positions from a global counter (the trial counts differ per trajectory)
struct Queue {
  unsigned long long* q;
  int lane;
  __device__ __forceinline__ int64_t next() const {
    unsigned long long nx = 0;
    if (lane == 0) nx = atomicAdd(q, 1ull);
    return (int64_t)__shfl_sync(FULL_MASK, nx, 0) + gridDim.x;
  }
};

template <class Opt>
__device__ __forceinline__ void load_states(Opt& o, const double* __restrict__ sp, bool with_dl) {
  const int n = o.N * Opt::b, lane = o.lane;
  for (int idx = lane; idx < n; idx += 32) o.xs[idx] = sp[idx];
  if (with_dl) {
    const double* dp = sp + pk_even(n);
    for (int idx = lane; idx < n; idx += 32) o.dl[idx] = dp[idx];
  }
}
template <class Opt>
__device__ __forceinline__ void set_ends(Opt& o, const KProblem& pr, int64_t prob) {
  constexpr int D = Opt::Dim;
  o.start_conf = pr.start_conf + prob * D; o.start_vel = pr.start_vel + prob * D;
  o.end_conf = pr.end_conf + prob * D;     o.end_vel = pr.end_vel + prob * D;
}

}  // namespace pk

// ------------------------------------------------------------------------------------------------------------------
// linearize kernel: M-list of every trajectory on the "needs linearization" list of this round
// ------------------------------------------------------------------------------------------------------------------
#ifndef PK_LIN_MIN_BLOCKS
#define PK_LIN_MIN_BLOCKS 1
#endif
template <class Opt>
__global__ void __launch_bounds__(32, PK_LIN_MIN_BLOCKS)
pk_lin_kernel(const __grid_constant__ KRobot rb, const __grid_constant__ KSdf sdf, const __grid_constant__ KSetting st,
              const __grid_constant__ KProblem pr, const double* __restrict__ hconst, int round) {
  extern __shared__ double smem[];
  constexpr int D = Opt::Dim;
  Opt o(rb, sdf, st, hconst, smem, false, 1);
  const int lane = o.lane, N = o.N, par = round & 1;
  int gather_chunk;   // the staging buffer doubles as the landing zone of the asynchronous SDF gathers: 2 buffers x chunk x 3 KB
  {
    unsigned dyn;
    asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
    gather_chunk = ((int)(dyn / sizeof(double)) - pk_even(N * Opt::b)) / (2 * 384);
  }
  const unsigned n = pr.pk_count[par * 2 + 0];
  // the lists of the NEXT round are appended to by this round's error kernel: reset their lengths here (their
  // previous contents were consumed by the previous round, which has completed)
  if (blockIdx.x == 0 && lane == 0) { pr.pk_count[(par ^ 1) * 2 + 0] = 0; pr.pk_count[(par ^ 1) * 2 + 1] = 0; }
  const int32_t* list = pk::list_of(pr, par, 0);
  const pk::Queue queue{pr.queue, lane};
  const int SS = pk_state_size(D, N), RS = pk_row_stride(D);
  const size_t MLS = pk_mlist_size(D, N, o.K);
  unsigned long long n_lin = 0;
  for (int64_t pos = blockIdx.x; pos < (int64_t)n; pos = queue.next()) {
    const int64_t prob = list[pos];
    DBG_IDX(prob, pr.B, "trajectory index from the work list");
    pk::load_states(o, pr.pk_state + prob * SS, false);
    __syncwarp();
    o.linearize_configs_to_global(pr.pk_mlist + prob * MLS, RS, smem + pk_even(N * Opt::b), gather_chunk);
    n_lin++;
    __syncwarp();
  }
  if (lane == 0 && pr.counters && n_lin) atomicAdd(pr.counters + 0, n_lin);
}

// ------------------------------------------------------------------------------------------------------------------
// linearize kernel of the Pose2Vector robots (LieOpt): the WHOLE linearization of the fused kernel (state-dependent
// GP-prior Hessian of GaussianProcessPriorLie, priors, limit hinges, obstacle factors through the block-sparse
// interpolation Jacobians of GaussianProcessInterpolatorLie) into H = [Ho | Hd] and g in shared memory, then one
// coalesced stream to the trajectory's H buffer in HBM (the layout pk_solve_mma_h_kernel loads: pk_hbuf_size).
// ------------------------------------------------------------------------------------------------------------------
#ifndef PK_LIE_LIN_MIN_BLOCKS
#define PK_LIE_LIN_MIN_BLOCKS 1
#endif
template <class Opt>
__global__ void __launch_bounds__(32, PK_LIE_LIN_MIN_BLOCKS)
pk_lin_full_kernel(const __grid_constant__ KRobot rb, const __grid_constant__ KSdf sdf, const __grid_constant__ KSetting st,
                   const __grid_constant__ KProblem pr, const double* __restrict__ hconst, int round) {
  extern __shared__ double smem[];
  constexpr int D = Opt::Dim, b = Opt::b;
  Opt o(rb, sdf, st, hconst, smem, true, 4);
  const int lane = o.lane, N = o.N, par = round & 1;
  const unsigned n = pr.pk_count[par * 2 + 0];
  if (blockIdx.x == 0 && lane == 0) { pr.pk_count[(par ^ 1) * 2 + 0] = 0; pr.pk_count[(par ^ 1) * 2 + 1] = 0; }
  const int32_t* list = pk::list_of(pr, par, 0);
  const pk::Queue queue{pr.queue, lane};
  const int SS = pk_state_size(D, N);
  const size_t HS = pk_hbuf_size(D, N);
  const int HSZ = (N - 1) * Opt::BB + N * Opt::BD, GOFF = pk_even(HSZ);
  unsigned long long n_lin = 0;
  for (int64_t pos = blockIdx.x; pos < (int64_t)n; pos = queue.next()) {
    const int64_t prob = list[pos];
    DBG_IDX(prob, pr.B, "trajectory index from the work list");
    pk::load_states(o, pr.pk_state + prob * SS, false);
    pk::set_ends(o, pr, prob);
    __syncwarp();
    o.mask_in = pr.pk_mask_use ? pr.pk_mask + prob * o.C : nullptr;
    o.template linearize<true>();
    double* H = pr.pk_mlist + prob * HS;
    {
      const double2* src = reinterpret_cast<const double2*>(o.Ho);
      double2* dst = reinterpret_cast<double2*>(H);
      for (int idx = lane; idx < (HSZ + 1) / 2; idx += 32) dst[idx] = src[idx];
      const double2* gs = reinterpret_cast<const double2*>(o.g);
      double2* gd = reinterpret_cast<double2*>(H + GOFF);
      for (int idx = lane; idx < (N * b) / 2; idx += 32) gd[idx] = gs[idx];
    }
    n_lin++;
    __syncwarp();
  }
  if (lane == 0 && pr.counters && n_lin) atomicAdd(pr.counters + 0, n_lin);
}

// ------------------------------------------------------------------------------------------------------------------
// solve kernel: H, g from the M-list; (H + lambda I) delta = -g; linearized cost change
// ------------------------------------------------------------------------------------------------------------------
template <class Opt>
__global__ void __launch_bounds__(32)
pk_solve_kernel(const __grid_constant__ KRobot rb, const __grid_constant__ KSdf sdf, const __grid_constant__ KSetting st,
                const __grid_constant__ KProblem pr, const double* __restrict__ hconst, int round) {
  extern __shared__ double smem[];
  constexpr int D = Opt::Dim, b = Opt::b;
#if PK_STREAMED_SOLVE
  Opt o(rb, sdf, st, hconst, smem, false, 3);
  double* slab = pr.h_backup + (size_t)blockIdx.x * pk_slab_size(D, st.N);
#else
  Opt o(rb, sdf, st, hconst, smem);
#endif
  const int lane = o.lane, N = o.N, par = round & 1;
  const unsigned n = pr.pk_count[par * 2 + 1];
  const int32_t* list = pk::list_of(pr, par, 1);
  const pk::Queue queue{pr.queue, lane};
  const int SS = pk_state_size(D, N), RS = pk_row_stride(D);
  const size_t MLS = pk_mlist_size(D, N, o.K);
  unsigned long long n_solve = 0;
  for (int64_t pos = blockIdx.x; pos < (int64_t)n; pos = queue.next()) {
    const int64_t prob = list[pos];
    double* sp = pr.pk_state + prob * SS;
    double* sc = sp + 2 * pk_even(N * b);
    pk::load_states(o, sp, false);
    pk::set_ends(o, pr, prob);
    const double lambda = sc[PKS_LAMBDA];
    __syncwarp();
    n_solve++;
#if PK_STREAMED_SOLVE
    o.assemble_diag_from_global(pr.pk_mlist + prob * MLS, RS);
    const bool solved = o.solve_streamed(lambda, pr.pk_mlist + prob * MLS, RS, slab);
#else
    o.assemble_from_global(pr.pk_mlist + prob * MLS, RS);
    const bool solved = o.solve(lambda);
#endif
    // linearized cost change = error - linear.error(delta) = -(g.delta) - 0.5 delta^T H delta
    //                        = -0.5 g.delta + 0.5 lambda |delta|^2   (using (H + lambda I) delta = -g)
    double gd = 0.0, dd = 0.0;
    for (int idx = lane; idx < N * b; idx += 32) {
      gd = fma(o.g[idx], o.dl[idx], gd);
      dd = fma(o.dl[idx], o.dl[idx], dd);
    }
    gd = warp_sum(gd);
    dd = warp_sum(dd);
    double* dp = sp + pk_even(N * b);
    for (int idx = lane; idx < N * b; idx += 32) dp[idx] = o.dl[idx];
    if (lane == 0) {
      sc[PKS_LIN_COST_CHANGE] = -0.5 * gd + 0.5 * lambda * dd;
      sc[PKS_SOLVED] = solved ? 1.0 : 0.0;
    }
    __syncwarp();
  }
  if (lane == 0 && pr.counters && n_solve) atomicAdd(pr.counters + 1, n_solve);
}

// ------------------------------------------------------------------------------------------------------------------
// error + decision kernel.  round < 0: the initial pass over ALL trajectories (load the wire trajectory, error at the
// initial values, lambda0); round >= 0: one tryLambda verdict for every trajectory on this round's solve list.
// ------------------------------------------------------------------------------------------------------------------
#ifndef PK_ERR_MIN_BLOCKS
#define PK_ERR_MIN_BLOCKS 20
#endif
template <class Opt>
__global__ void __launch_bounds__(32, PK_ERR_MIN_BLOCKS)
pk_err_kernel(const __grid_constant__ KRobot rb, const __grid_constant__ KSdf sdf, const __grid_constant__ KSetting st,
              const __grid_constant__ KProblem pr, const double* __restrict__ hconst, int round) {
  extern __shared__ double smem[];
  constexpr int D = Opt::Dim, b = Opt::b;
  Opt o(rb, sdf, st, hconst, smem, false, 2);
  const int lane = o.lane, N = o.N;
  {   // whatever dynamic shared memory the host gave beyond xs | dl is the landing zone of the asynchronous SDF gathers
    unsigned dyn;
    asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
    const int base = Opt::LIE ? pk_lie_err_smem(D, N) : pk_small_smem(D, N, true);
    o.escr = smem + base;
    o.escr_chunk = ((int)(dyn / sizeof(double)) - base) / 384;
  }
  const bool init = round < 0;
  const int par = round & 1;                    // (init: round = -1 -> the lists of parity 0 are the ones to fill)
  const int wpar = init ? 0 : par ^ 1;
  const int64_t n = init ? pr.B : (int64_t)pr.pk_count[par * 2 + 1];
  const int32_t* list = pk::list_of(pr, par, 1);
  int32_t* next_lin = pk::list_of_mut(pr, wpar, 0);
  int32_t* next_solve = pk::list_of_mut(pr, wpar, 1);
  const pk::Queue queue{pr.queue, lane};
  const int SS = pk_state_size(D, N), TL = 2 * N * D;
  const double lambdaFactor = 10.0, lambdaUpperBound = 1e5, lambdaLowerBound = 0.0, minModelFidelity = 1e-3;
  const double absoluteErrorTol = 1e-5, errorTol = 0.0, relativeErrorTol = st.rel_thresh;
  unsigned long long n_err = 0;
  for (int64_t pos = blockIdx.x; pos < n; pos = queue.next()) {
    const int64_t prob = init ? pos : (int64_t)list[pos];
    DBG_IDX(prob, pr.B, "trajectory index from the work list");
    double* sp = pr.pk_state + prob * SS;
    double* sc = sp + 2 * pk_even(N * b);
    pk::set_ends(o, pr, prob);
    // every error evaluation records the sphere masks for the linearization that may follow it (config_eval<MASKED>);
    // one instantiation of the error pass only -- the kernel runs at 96 registers
    o.mask_out = pr.pk_mask + prob * o.C;
    double lambda, error, currentError;
    int iterations, status;
    bool finished = false, relinearize = false;
    if (init) {
      // ---- load the trajectory (wire layout [x_0..x_T | v_0..v_T]); error at xs = error at xs + 0 ----
      const double* tin = pr.init_traj + prob * TL;
      for (int idx = lane; idx < N * D; idx += 32) {
        const int i = idx / D, d = idx - i * D;
        o.xs[i * b + d] = tin[idx];
        o.xs[i * b + D + d] = tin[N * D + idx];
      }
      for (int idx = lane; idx < N * b; idx += 32) o.dl[idx] = 0.0;
      __syncwarp();
      lambda = 100.0;                            // setlambdaInitial(100.0), BatchTrajOptimizer.cpp:226
      // (Pose2Vector states: retract(x, 0) re-wraps theta, so evaluate at xs itself)
      if constexpr (Opt::LIE) error = o.template eval_error<false, true>();
      else error = o.template eval_error<true, true>();
      n_err++;
      currentError = error;
      iterations = 0; status = 0;
      for (int idx = lane; idx < N * b; idx += 32) sp[idx] = o.xs[idx];
      if (currentError <= errorTol) { status |= 32; finished = true; }
      else if (iterations >= st.max_iter) { status |= 4; finished = true; }
      else relinearize = true;
    } else {
      pk::load_states(o, sp, true);
      lambda = sc[PKS_LAMBDA]; error = sc[PKS_ERROR]; currentError = sc[PKS_CURRENT_ERROR];
      iterations = (int)sc[PKS_ITERATIONS]; status = (int)sc[PKS_STATUS];
      const bool solved = sc[PKS_SOLVED] != 0.0;
      const double linearizedCostChange = sc[PKS_LIN_COST_CHANGE];
      __syncwarp();
      // ---- LevenbergMarquardtOptimizer::tryLambda after the solve [GTSAM semantics, SURVEY.md App. B.1] ----
      bool step_is_successful = false, stopSearchingLambda = false;
      double newError = 0.0;
      if (solved) {
        if (linearizedCostChange >= 0.0) {
          newError = o.template eval_error<true, true>();
          n_err++;
          const double costChange = error - newError;
          if (linearizedCostChange > 2.220446049250313e-16 * fabs(error)) {
            const double modelFidelity = costChange / linearizedCostChange;
            step_is_successful = modelFidelity > minModelFidelity;
          }
          const double minAbsoluteTolerance = relativeErrorTol * error;
          if (fabs(costChange) < minAbsoluteTolerance) stopSearchingLambda = true;
        }
      } else {
        status |= 16;
      }
      bool iterate_done = true;                  // this iterate() of gpmp2::optimize's loop is over
      if (step_is_successful) {
        o.accept_step();
        for (int idx = lane; idx < N * b; idx += 32) sp[idx] = o.xs[idx];
        error = newError;
        lambda = fmax(lambdaLowerBound, lambda / lambdaFactor);
        iterations++;
      } else if (!stopSearchingLambda) {
        lambda *= lambdaFactor;
        if (lambda >= lambdaUpperBound) status |= 8;     // "giving up": the iterate ends without a step
        else iterate_done = false;                       // retry with the larger lambda: solve only
      }
      if (iterate_done) {
        // checkConvergence(relativeErrorTol, absoluteErrorTol, errorTol, currentError, error)  [App. B.2]
        bool converged = false;
        int why = 0;
        if (error <= errorTol) { converged = true; why = 32; }
        else {
          const double absoluteDecrease = currentError - error;
          const double relativeDecrease = absoluteDecrease / currentError;
          const bool rel = (relativeErrorTol != 0.0) && (relativeDecrease <= relativeErrorTol);
          const bool ab = absoluteDecrease <= absoluteErrorTol;
          converged = rel || ab;
          why = (rel ? 2 : 0) | (ab ? 1 : 0);
        }
        if (iterations < st.max_iter && !converged) {
          relinearize = true; currentError = error;
          // the stored masks belong to the states of the last error evaluation: if that was not an accepted step (cannot
          // happen -- an iterate without a step has zero decrease and converges -- but cheap to guard), lift them
          if (!step_is_successful) { for (int c = lane; c < o.C; c += 32) o.mask_out[c] = ~0ull; }
        }
        else {
          finished = true;
          if (iterations >= st.max_iter) status |= 4;
          else status |= why;
          // (BatchTrajOptimizer.cpp:297-307, "return last_values if the error increased": an accepted LM step never
          //  increases the error and a rejected one leaves the values alone, so xs already is the answer)
        }
      }
    }
    if (finished) {
      double* tout = pr.out_traj + prob * TL;
      for (int idx = lane; idx < N * D; idx += 32) {
        const int i = idx / D, d = idx - i * D;
        tout[idx] = o.xs[i * b + d];
        tout[N * D + idx] = o.xs[i * b + D + d];
      }
      if (lane == 0) {
        if (pr.out_error) pr.out_error[prob] = error;
        if (pr.out_iters) pr.out_iters[prob] = iterations;
        if (pr.out_status) pr.out_status[prob] = status;
      }
    } else if (lane == 0) {
      sc[PKS_LAMBDA] = lambda; sc[PKS_ERROR] = error; sc[PKS_CURRENT_ERROR] = currentError;
      sc[PKS_ITERATIONS] = (double)iterations; sc[PKS_STATUS] = (double)status;
      if (relinearize) { const unsigned k = atomicAdd(pr.pk_count + wpar * 2 + 0, 1u); DBG_IDX(k, pr.B, "work list slot"); next_lin[k] = (int32_t)prob; }
      { const unsigned k = atomicAdd(pr.pk_count + wpar * 2 + 1, 1u); DBG_IDX(k, pr.B, "work list slot"); next_solve[k] = (int32_t)prob; }
    }
    __syncwarp();
  }
  if (lane == 0 && pr.counters && n_err) atomicAdd(pr.counters + 2, n_err);
}
