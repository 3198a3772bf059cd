"""Host-side mirror of the reference's BA interface over the C ABI.

`BundleAdjustmentOptions` has the reference's five fields with the same
defaults (include/visnav/map_utils.h:304-319); `bundle_adjustment(problem,
options)` is the drop-in call (map_utils.h:322): host buffers in, poses /
inverse distances (/ affine) updated in place, a Ceres-style summary returned.
`Engine` exposes the split, device-resident entry points used by the parity
tests and bench.py.
"""
import ctypes as C
import dataclasses

import numpy as np

from . import _ffi
from .problem import Problem


@dataclasses.dataclass
class BundleAdjustmentOptions:
    # -- the reference's fields (map_utils.h:304-319) --
    verbosity_level: int = 1
    optimize_intrinsics: bool = False
    use_huber: bool = True
    huber_parameter: float = 1.0
    max_num_iterations: int = 20
    # -- engine extensions (defaults = Ceres 2.0.0 Solver::Options) --
    solver: int = _ffi.SOLVER_AUTO
    cholesky_max_dim: int = None
    pcg_max_iterations: int = None
    pcg_tolerance: float = None
    function_tolerance: float = None
    gradient_tolerance: float = None
    parameter_tolerance: float = None
    initial_trust_region_radius: float = None
    device: int = 0
    num_gpus: int = None  # pba_solve from one process over this many GPUs (None = 1, 0 = all visible)
    profile: int = 0  # 0 none, 1 CUDA events around every kernel, 2 only around the residual/Jacobian kernel

    def to_c(self):
        o = _ffi.pba_options()
        _ffi.load_lib().pba_options_init(C.byref(o))
        for f in dataclasses.fields(self):
            v = getattr(self, f.name)
            if v is None:
                continue
            setattr(o, f.name, int(v) if isinstance(v, bool) else v)
        return o


class Summary:
    """ceres::Solver::Summary look-alike filled from pba_summary."""

    def __init__(self, max_iterations=64):
        self._its = (_ffi.pba_iteration * (max_iterations + 2))()
        self.c = _ffi.pba_summary()
        self.c.iterations = C.cast(self._its, C.POINTER(_ffi.pba_iteration))
        self.c.iterations_capacity = max_iterations + 2

    def __getattr__(self, name):
        c = object.__getattribute__(self, "c")
        if name == "message":
            return c.message.decode()
        return getattr(c, name)

    @property
    def iterations(self):
        out = []
        for i in range(min(self.c.num_iterations, self.c.iterations_capacity)):
            it = self._its[i]
            out.append({f[0]: getattr(it, f[0]) for f in _ffi.pba_iteration._fields_})
        return out

    def brief_report(self):
        term = {0: "CONVERGENCE", 1: "NO_CONVERGENCE", 2: "FAILURE"}.get(self.c.termination_type, "?")
        return ("B200 PBA Report: Iterations: %d, Initial cost: %e, Final cost: %e, Termination: %s"
                % (self.c.num_iterations, self.c.initial_cost, self.c.final_cost, term))


def device_count():
    return int(_ffi.load_lib().pba_device_count())


def analyze_structure(problem: Problem, options: BundleAdjustmentOptions = None):
    """Host-only: (slot per pose, n_slots, natural half-bandwidth, final half-bandwidth, stored RCS blocks)."""
    o = (options or BundleAdjustmentOptions()).to_c()
    slot = np.zeros(problem.n_poses, np.int32)
    ns, bw0, bw1, nb = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int64()
    pc = problem.c
    _ffi.check(_ffi.load_lib().pba_analyze_structure(C.byref(pc), C.byref(o), _ffi.ptr(slot, C.c_int32), C.byref(ns),
                                                     C.byref(bw0), C.byref(bw1), C.byref(nb)), "pba_analyze_structure")
    return slot, ns.value, bw0.value, bw1.value, nb.value


def multi_gpu_init(device=0, num_gpus=0):
    """Create the NCCL communicators of bundle_adjustment(..., num_gpus > 1) ahead of the first solve."""
    _ffi.check(_ffi.load_lib().pba_multi_gpu_init(int(device), int(num_gpus)), "pba_multi_gpu_init")


def bundle_adjustment(problem: Problem, options: BundleAdjustmentOptions = None) -> Summary:
    """Drop-in for visnav::bundle_adjustment (map_utils.h:322-399) on flat containers."""
    options = options or BundleAdjustmentOptions()
    lib = _ffi.load_lib()
    s = Summary(options.max_num_iterations)
    o = options.to_c()
    pc = problem.c
    _ffi.check(lib.pba_solve(C.byref(pc), C.byref(o), C.byref(s.c)), "pba_solve")
    return s


class Engine:
    """Device-resident problem: the split entry points of include/pba.h."""

    def __init__(self, problem: Problem, options: BundleAdjustmentOptions = None, rank=0, world_size=1):
        self.lib = _ffi.load_lib()
        self.problem = problem
        self.options = options or BundleAdjustmentOptions()
        self._o = self.options.to_c()
        self._h = C.c_void_p()
        pc = problem.c
        _ffi.check(self.lib.pba_create(C.byref(pc), C.byref(self._o), rank, world_size, C.byref(self._h)),
                   "pba_create")
        n_obs, n_lm, first = C.c_int64(), C.c_int32(), C.c_int64()
        _ffi.check(self.lib.pba_get_sizes(self._h, C.byref(n_obs), C.byref(n_lm), C.byref(first)))
        self.n_obs_local, self.n_landmarks_local, self.first_landmark = n_obs.value, n_lm.value, first.value

    def close(self):
        if self._h:
            self.lib.pba_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def set_stream(self, cuda_stream_ptr):
        _ffi.check(self.lib.pba_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def synchronize(self):
        _ffi.check(self.lib.pba_synchronize(self._h))

    def evaluate(self, with_jacobian=True, want_cost=True):
        cost = C.c_double()
        _ffi.check(self.lib.pba_evaluate(self._h, int(with_jacobian), C.byref(cost) if want_cost else None),
                   "pba_evaluate")
        return cost.value if want_cost else None

    def residuals(self):
        r = np.zeros((self.n_obs_local, self.problem.res_per_obs))
        _ffi.check(self.lib.pba_get_residuals(self._h, _ffi.ptr(r, C.c_double)))
        return r

    def jacobians(self):
        J = np.zeros((self.n_obs_local, self.problem.res_per_obs, self.problem.cols_per_obs))
        _ffi.check(self.lib.pba_get_jacobians(self._h, _ffi.ptr(J, C.c_double)))
        return J

    def blocks(self, obs_index, want_jacobians=True):
        """Residuals [n,R] and local Jacobians [n,R,C] of the selected blocks (caller-order local indices)."""
        idx = np.ascontiguousarray(obs_index, np.int64)
        r = np.zeros((idx.shape[0], self.problem.res_per_obs))
        J = np.zeros((idx.shape[0], self.problem.res_per_obs, self.problem.cols_per_obs)) if want_jacobians else None
        _ffi.check(self.lib.pba_get_blocks(self._h, idx.shape[0], _ffi.ptr(idx, C.c_int64), _ffi.ptr(r, C.c_double),
                                           _ffi.ptr(J, C.c_double)), "pba_get_blocks")
        return r, J

    def build_rcs(self, radius=1e4):
        _ffi.check(self.lib.pba_build_rcs(self._h, float(radius)), "pba_build_rcs")

    def rcs(self):
        d = C.c_int32()
        _ffi.check(self.lib.pba_get_rcs_dim(self._h, C.byref(d)))
        S = np.zeros((d.value, d.value))
        rhs = np.zeros(d.value)
        _ffi.check(self.lib.pba_get_rcs(self._h, _ffi.ptr(S, C.c_double), _ffi.ptr(rhs, C.c_double)))
        return S, rhs

    def solve_rcs(self, solver=_ffi.SOLVER_CHOLESKY):
        d = C.c_int32()
        _ffi.check(self.lib.pba_get_rcs_dim(self._h, C.byref(d)))
        y = np.zeros(d.value)
        it = C.c_int32()
        _ffi.check(self.lib.pba_solve_rcs(self._h, solver, _ffi.ptr(y, C.c_double), C.byref(it)), "pba_solve_rcs")
        return y, it.value

    def minimize(self):
        s = Summary(self.options.max_num_iterations)
        _ffi.check(self.lib.pba_minimize(self._h, C.byref(s.c)), "pba_minimize")
        return s

    def lm_iterate(self, radius=1e4, apply=False):
        it = _ffi.pba_iteration()
        _ffi.check(self.lib.pba_lm_iterate(self._h, float(radius), int(apply), C.byref(it)), "pba_lm_iterate")
        return {f[0]: getattr(it, f[0]) for f in _ffi.pba_iteration._fields_}

    def set_state(self, poses, inv_depth, affine=None):
        poses = np.ascontiguousarray(poses, np.float64)
        inv_depth = np.ascontiguousarray(inv_depth, np.float64)
        affine = None if affine is None else np.ascontiguousarray(affine, np.float64)
        _ffi.check(self.lib.pba_set_state(self._h, _ffi.ptr(poses, C.c_double), _ffi.ptr(inv_depth, C.c_double),
                                          _ffi.ptr(affine, C.c_double)))

    def get_state(self):
        poses = np.zeros((self.problem.n_poses, 7))
        rho = np.zeros(self.n_landmarks_local)
        aff = np.zeros((self.problem.n_poses, 2)) if self.problem.mode == _ffi.MODE_PHOTOMETRIC else None
        _ffi.check(self.lib.pba_get_state(self._h, _ffi.ptr(poses, C.c_double), _ffi.ptr(rho, C.c_double),
                                          _ffi.ptr(aff, C.c_double)))
        return poses, rho, aff

    def reset_kernel_stats(self):
        _ffi.check(self.lib.pba_reset_kernel_stats(self._h))

    def set_profile(self, level):
        """0 none, 1 CUDA events around every kernel, 2 only around the residual/Jacobian kernel."""
        _ffi.check(self.lib.pba_set_profile(self._h, int(level)))

    def kernel_stats(self):
        buf = (_ffi.pba_kernel_stat * 64)()
        n = self.lib.pba_get_kernel_stats(self._h, buf, 64)
        return {buf[i].name.decode(): (buf[i].launches, buf[i].total_ms) for i in range(n)}

    @property
    def collective(self):
        """'none' | 'nccl' | 'peer' (include/pba.h: pba_collective_kind)."""
        return {0: "none", 1: "nccl", 2: "peer"}[int(self.lib.pba_collective_kind(self._h))]

    def comm_init(self, nccl_id: bytes):
        arr = (C.c_uint8 * _ffi.NCCL_ID_BYTES).from_buffer_copy(nccl_id)
        _ffi.check(self.lib.pba_comm_init(self._h, arr), "pba_comm_init")


def nccl_unique_id() -> bytes:
    arr = (C.c_uint8 * _ffi.NCCL_ID_BYTES)()
    _ffi.check(_ffi.load_lib().pba_nccl_unique_id(arr), "pba_nccl_unique_id")
    return bytes(arr)
