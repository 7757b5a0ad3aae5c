// kernels_inst.cu -- one translation unit per (robot kind, system dof): compile with -DINST_KIND=vec|lie -DINST_D=<D>.
// Instantiates, for both SDF dimensions, the three optimizer kernels (Gauss-Newton, LM, Dogleg) and the auxiliary
// kernel, and exports the lookup c_abi.cu dispatches through.  Separate optimizer instantiations keep each
// kernel's hot code small (the kernel is instruction-cache sensitive).
#include "optimizer_kernel_lie.cuh"
#include "pk_kernels.cuh"
#include "pk_solve_mma.cuh"

#define CAT_(a, b, c) a##b##_##c
#define CAT(a, b, c) CAT_(a, b, c)

#if INST_IS_LIE
template <int NDIM> using OptT = LieOpt<INST_D, NDIM>;
template <int NDIM> using GoalOptT = LieOpt<INST_D, NDIM, true>;   // + optional workspace-goal / self-collision factors
#define LOOKUP CAT(gpmp2b_lookup_, lie, INST_D)
#else
template <int NDIM> using OptT = VecOpt<INST_D, NDIM>;
template <int NDIM> using GoalOptT = VecOpt<INST_D, NDIM, true>;
#define LOOKUP CAT(gpmp2b_lookup_, vec, INST_D)
#endif

template <int NDIM>
static KernelFn pick(int opt) {
  switch (opt) {
    case 0: return gpmp2b_kernel<OptT<NDIM>, 0>;
    case 1: return gpmp2b_kernel<OptT<NDIM>, 1>;
    case 2: return gpmp2b_kernel<OptT<NDIM>, 2>;
    case -1: return gpmp2b_kernel<OptT<NDIM>, -1>;
    case KOPT_GOAL + 0: return gpmp2b_kernel<GoalOptT<NDIM>, 0>;
    case KOPT_GOAL + 1: return gpmp2b_kernel<GoalOptT<NDIM>, 1>;
    case KOPT_GOAL + 2: return gpmp2b_kernel<GoalOptT<NDIM>, 2>;
    case KOPT_GOAL - 1: return gpmp2b_kernel<GoalOptT<NDIM>, -1>;
#if INST_IS_LIE
    // phase-kernel pipeline of the LM optimizer, Pose2Vector states: full linearization -> H in HBM, tensor-core solve
    case KOPT_PK_LINH: return pk_lin_full_kernel<OptT<NDIM>>;
    case KOPT_PK_ERR: return pk_err_kernel<OptT<NDIM>>;
    case KOPT_PK_SOLVE_MMA_H: return pk_solve_mma_h_kernel<INST_D>;
#endif
#if !INST_IS_LIE
    // phase-kernel pipeline of the LM optimizer (pk_kernels.cuh)
    case KOPT_PK_LIN: return pk_lin_kernel<OptT<NDIM>>;
    case KOPT_PK_SOLVE: return pk_solve_kernel<OptT<NDIM>>;
    case KOPT_PK_ERR: return pk_err_kernel<OptT<NDIM>>;
    case KOPT_PK_SOLVE_MMA: return pk_solve_mma_kernel<INST_D>;
    case KOPT_PK_LINH: return pk_linh_kernel<OptT<NDIM>>;
    case KOPT_PK_SOLVE_MMA_H: return pk_solve_mma_h_kernel<INST_D>;
#endif
  }
  return nullptr;
}

KernelFn LOOKUP(int ndim, int opt) { return ndim == 3 ? pick<3>(opt) : (ndim == 2 ? pick<2>(opt) : nullptr); }
