"""ctypes binding of libmpb_b200.so (the C ABI declared in include/mpb.h).

There is NO fallback: if the shared library is missing or a call fails, we raise.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libmpb_b200.so')

MPB_MAX_FIELDS = 4
ROBOT_POINT, ROBOT_CHAIN = 0, 1


class MpbError(RuntimeError):
    pass


class RobotDesc(C.Structure):
    _fields_ = [('kind', C.c_int32), ('q_dim', C.c_int32), ('ws_dim', C.c_int32), ('n_spheres', C.c_int32),
                ('fixed_tf', C.c_void_p), ('sphere_link', C.c_void_p), ('sphere_off', C.c_void_p),
                ('sphere_r', C.c_void_p)]


FIELD_PRIMITIVES, FIELD_SELF, FIELD_WORKSPACE = 0, 1, 2
MPB_MAX_SELF_PAIRS = 4096


class FieldDesc(C.Structure):
    _fields_ = [('n_spheres', C.c_int32), ('n_boxes', C.c_int32), ('spheres', C.c_void_p), ('boxes', C.c_void_p),
                ('cutoff_margin', C.c_float), ('weight', C.c_float), ('inv_sigma2', C.c_float),
                ('kind', C.c_int32), ('n_pairs', C.c_int32), ('pairs', C.c_void_p),
                ('ws_min', C.c_float * 3), ('ws_max', C.c_float * 3)]


class GPDesc(C.Structure):
    _fields_ = [('enabled', C.c_int32), ('has_goal', C.c_int32), ('dt', C.c_float), ('k_start', C.c_float),
                ('k_goal', C.c_float), ('q11', C.c_float), ('q12', C.c_float), ('q22', C.c_float),
                ('w_gp', C.c_float), ('w_goal', C.c_float), ('start_state', C.c_void_p), ('goal_state', C.c_void_p)]


class ExtraCostDesc(C.Structure):
    _fields_ = [('gp_traj_enabled', C.c_int32), ('t11', C.c_float), ('t12', C.c_float), ('t22', C.c_float),
                ('w_gp_traj', C.c_float), ('jl_enabled', C.c_int32), ('jl_eps', C.c_float), ('w_jl', C.c_float),
                ('q_min', C.c_void_p), ('q_max', C.c_void_p)]


class NoiseDesc(C.Structure):
    """mpb_noise_desc: in-kernel Philox noise keyed on the global element index (include/mpb.h)."""
    _fields_ = [('seed', C.c_uint64), ('offset', C.c_uint64), ('s_offset', C.c_int64), ('p_offset', C.c_int64),
                ('P_global', C.c_int64)]


NOISE_SPM, NOISE_STOMP, NOISE_MPPI, NOISE_SPMD = 0, 1, 2, 3
MPB_MAX_INTERP = 32
_lib = None

_vp, _i, _f = C.c_void_p, C.c_int, C.c_float
_SIGNATURES = {
    'mpb_last_error': (C.c_char_p, []),
    'mpb_version': (C.c_int, []),
    'mpb_init': (C.c_int, []),
    'mpb_bench_fp32_peak': (C.c_int, [_vp, _i, C.POINTER(C.c_longlong), _vp]),
    'mpb_sizeof_desc': (C.c_int, [_i]),
    'mpb_sample_gp': (C.c_int, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    'mpb_split_tf32': (C.c_int, [_vp, _vp, _vp, C.c_longlong, _vp]),
    'mpb_sample_gp_tc_supported': (C.c_int, [_i, _i, _i]),
    'mpb_sample_gp_tc': (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    'mpb_sample_gp_kron_supported': (C.c_int, [_i, _i]),
    'mpb_sample_gp_kron_pack': (C.c_int, [_vp, _vp, _i, _i, C.POINTER(C.c_int), _vp]),
    'mpb_sample_gp_kron': (C.c_int, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    'mpb_sample_gp_kron_tc_bytes': (C.c_longlong, [_i, _i]),
    'mpb_sample_gp_kron_tc_prepare': (C.c_int, [_vp, _vp, _i, _i, _vp]),
    'mpb_sample_gp_kron_tc': (C.c_int, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    'mpb_sample_stomp': (C.c_int, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    'mpb_sample_stomp_rng': (C.c_int, [_vp, _vp, C.POINTER(NoiseDesc), _vp, _i, _i, _i, _i, _vp]),
    'mpb_philox_normal': (C.c_int, [C.POINTER(NoiseDesc), _i, _vp, _i, _i, _i, _i, _vp]),
    'mpb_sample_gp_kron_tc_rng': (C.c_int, [_vp, _vp, C.POINTER(NoiseDesc), _vp, _i, _i, _i, _i, _vp]),
    'mpb_stoch_gpmp_iter_kron_rng': (C.c_int, [_vp, _vp, _i, C.POINTER(NoiseDesc), _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i,
                                               C.POINTER(RobotDesc), C.POINTER(FieldDesc), _i, C.POINTER(GPDesc), _f, _f, _vp]),
    'mpb_stoch_gpmp_iter_kron_gen': (C.c_int, [_vp, _vp, _i, C.POINTER(NoiseDesc), _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i,
                                               C.POINTER(RobotDesc), C.POINTER(FieldDesc), _i, C.POINTER(GPDesc), _f, _f, _vp]),
    'mpb_stoch_gpmp_iter_kron_gen_ex': (C.c_int, [_vp, _vp, _i, C.POINTER(NoiseDesc), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i,
                                                  C.POINTER(RobotDesc), C.POINTER(FieldDesc), _i, C.POINTER(GPDesc), _f, _f, _vp]),
    'mpb_stomp_run': (C.c_int, [_vp, _vp, C.POINTER(NoiseDesc), _vp, _vp, _vp, _vp, _i, _i, _i, C.POINTER(RobotDesc),
                                C.POINTER(FieldDesc), _i, C.POINTER(GPDesc), _f, _f, _i, _vp]),
    'mpb_sample_gp_kron_gen_supported': (C.c_int, [_i, _i]),
    'mpb_sample_gp_kron_gen_bytes': (C.c_longlong, [_i, _i]),
    'mpb_sample_gp_kron_gen_prepare': (C.c_int, [_vp, _vp, _i, _i, _vp]),
    'mpb_sample_gp_kron_gen': (C.c_int, [_vp, _vp, C.POINTER(NoiseDesc), _vp, _i, _i, _i, _i, _vp]),
    'mpb_sample_gp_kron_gen_mv': (C.c_int, [_vp, _vp, C.POINTER(NoiseDesc), _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    'mpb_sample_gp_kron_gen_dm': (C.c_int, [_vp, _vp, C.POINTER(NoiseDesc), _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    'mpb_cost_eval_dm_supported': (C.c_int, [C.POINTER(RobotDesc), C.POINTER(FieldDesc), _i, _i]),
    'mpb_cost_eval_dm': (C.c_int, [_vp, _i, _i, C.POINTER(RobotDesc), C.POINTER(FieldDesc), _i, C.POINTER(GPDesc),
                                   _vp, _i, _f, _vp, _vp, _vp, _vp]),
    'mpb_softmax_update_dm': (C.c_int, [_vp, _vp, _vp, _vp, _vp, _f, _f, _vp, _i, _i, _i, _i, _vp]),
    'mpb_traj_from_dof_major': (C.c_int, [_vp, _vp, C.c_longlong, _i, _i, _vp]),
    'mpb_traj_to_dof_major': (C.c_int, [_vp, _vp, C.c_longlong, _i, _i, _vp]),
    'mpb_stoch_gpmp_iter_kron_gen_dm': (C.c_int, [_vp, _vp, C.POINTER(NoiseDesc), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i,
                                                  C.POINTER(RobotDesc), C.POINTER(FieldDesc), _i, C.POINTER(GPDesc), _f, _f, _vp]),
    'mpb_mppi_rollout_ex': (C.c_int, [_vp] * 5 + [C.POINTER(NoiseDesc)] + [_vp] * 7 + [_i, _i, _i, _i, _f, _f, _f, _f, _f, _vp]),
    'mpb_mppi_rollout_opt': (C.c_int, [_vp] * 5 + [C.POINTER(NoiseDesc)] + [_vp] * 7 + [_i, _i, _i, _i, _f, _f, _f, _f, _f, _i, _vp]),
    'mpb_prior_matvec': (C.c_int, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    'mpb_cost_eval': (C.c_int, [_vp, _i, _i, C.POINTER(RobotDesc), C.POINTER(FieldDesc), _i, C.POINTER(GPDesc),
                                _vp, _i, _f, _vp, _vp, _vp, _vp]),
    'mpb_cost_eval_ex': (C.c_int, [_vp, _i, _i, C.POINTER(RobotDesc), C.POINTER(FieldDesc), _i, C.POINTER(GPDesc),
                                   _vp, _i, _f, _vp, _vp, _vp, C.POINTER(ExtraCostDesc), _vp, _vp]),
    'mpb_smoothness_cost': (C.c_int, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    'mpb_softmax_update': (C.c_int, [_vp, _vp, _vp, _vp, _vp, _f, _f, _vp, _i, _i, _i, _i, _vp]),
    'mpb_softmax_update_ex': (C.c_int, [_vp, _vp, _vp, _vp, _vp, _f, _f, _vp, _vp, _i, _i, _i, _i, _vp]),
    'mpb_stoch_gpmp_iter': (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i,
                                      C.POINTER(RobotDesc), C.POINTER(FieldDesc), _i, C.POINTER(GPDesc),
                                      _f, _f, _vp]),
    'mpb_prior_dof_structured': (C.c_int, [_vp, _i, _i, C.POINTER(C.c_int), _vp]),
    'mpb_prior_matvec_dof': (C.c_int, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    'mpb_sample_gp_kron_umma_supported': (C.c_int, [_i, _i]),
    'mpb_sample_gp_kron_umma_floats': (C.c_longlong, [_i, _i]),
    'mpb_sample_gp_kron_umma_prepare': (C.c_int, [_vp, _vp, _i, _i, _vp]),
    'mpb_sample_gp_kron_umma': (C.c_int, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    'mpb_stoch_gpmp_iter_kron': (C.c_int, [_vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i,
                                           C.POINTER(RobotDesc), C.POINTER(FieldDesc), _i, C.POINTER(GPDesc),
                                           _f, _f, _vp]),
    'mpb_chomp_run': (C.c_int, [_vp, _i, _i, C.POINTER(RobotDesc), C.POINTER(FieldDesc), _i, _vp, _f, _f, _f, _i, _vp]),
    'mpb_chomp_run_ex': (C.c_int, [_vp, _i, _i, C.POINTER(RobotDesc), C.POINTER(FieldDesc), _i, _vp, _f, _f, _f, _i,
                                   C.POINTER(ExtraCostDesc), _vp]),
    'mpb_gpmp2_linearize_ex': (C.c_int, [_vp, _i, _i, C.POINTER(RobotDesc), C.POINTER(FieldDesc), _i, _vp, _vp, _vp,
                                         _i, C.POINTER(C.c_float), _vp]),
    'mpb_gpmp2_linearize': (C.c_int, [_vp, _i, _i, C.POINTER(RobotDesc), C.POINTER(FieldDesc), _i, _vp, _vp, _vp, _vp]),
    'mpb_gpmp2_workspace_bytes': (C.c_longlong, [_i, _i, _i]),
    'mpb_gpmp2_solve': (C.c_int, [_vp, _i, _i, _i, C.POINTER(GPDesc), _vp, _vp, C.POINTER(C.c_float), _i, _vp, _f, _f,
                                  _vp, _vp, _vp, _vp]),
    'mpb_softmax_record_len': (C.c_int, [_i, _i]),
    'mpb_softmax_partial': (C.c_int, [_vp, _vp, _vp, _vp, _f, _i, _i, _i, _i, _i, _i, _i, C.c_longlong, _vp]),
    'mpb_softmax_combine': (C.c_int, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _f, _vp, _i, _i, _i, _vp]),
    'mpb_softmax_weights': (C.c_int, [_vp, _vp, _vp, _f, _i, _i, _vp]),
    'mpb_sum_f64': (C.c_int, [_vp, C.c_longlong, _vp, _vp, _vp]),
    'mpb_mppi_rollout': (C.c_int, [_vp] * 11 + [_i, _i, _i, _i, _f, _f, _f, _f, _f, _vp]),
    'mpb_mppi_finalize': (C.c_int, [_vp, _vp, _vp, _f, _vp, _i, _i, _vp]),
    'mpb_fk_spheres': (C.c_int, [_vp, C.c_longlong, C.POINTER(RobotDesc), _vp, _vp]),
    'mpb_fk_spheres_vjp': (C.c_int, [_vp, _vp, C.c_longlong, C.POINTER(RobotDesc), _vp, _vp]),
    'mpb_field_cost': (C.c_int, [_vp, C.c_longlong, C.POINTER(RobotDesc), C.POINTER(FieldDesc), _vp, _vp, _vp]),
    'mpb_cost_grad': (C.c_int, [_vp, _i, _i, C.POINTER(RobotDesc), C.POINTER(FieldDesc), _i, C.POINTER(GPDesc),
                                C.POINTER(ExtraCostDesc), _vp, _vp, _vp, _vp]),
    'mpb_collision_query': (C.c_int, [_vp, C.c_longlong, _i, C.POINTER(RobotDesc), C.POINTER(FieldDesc), _i, _vp, _vp, _vp]),
}


def exported_symbols():
    """Every symbol include/mpb.h declares (checked by the CPU test-suite)."""
    return sorted(_SIGNATURES)


class _StreamOfCall:
    """Placeholder returned by stream_ptr(): resolved by the call wrapper to the CURRENT stream of the device the
    call's tensors live on (not of whatever device happens to be current)."""


_STREAM = _StreamOfCall()


class DevPtr(C.c_void_p):
    """c_void_p that remembers which CUDA device the tensor lives on (checked by the call wrapper)."""
    mpb_device = None


class _Fn:
    """One C-ABI entry point.  Every tensor argument of a call must live on ONE device; the call runs with that
    device current and on that device's current stream -- the library itself sizes grids and takes its scheduler
    slot from cudaGetDevice() -- so a planner built on cuda:1 works while cuda:0 is current, as torch ops would."""
    __slots__ = ('fn', 'name')

    def __init__(self, fn, name):
        self.fn, self.name = fn, name

    def __call__(self, *args):
        dev = None
        has_stream = False
        for a in args:
            if a is _STREAM:
                has_stream = True
                continue
            d = getattr(a, 'mpb_device', None)
            if d is None:
                continue
            if dev is None:
                dev = d
            elif d != dev:
                raise MpbError(f'{self.name}: tensors of one call live on different devices (cuda:{dev} and cuda:{d})')
        if dev is None:
            dev = torch.cuda.current_device() if has_stream else None
        if has_stream:
            sp = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            args = tuple(sp if a is _STREAM else a for a in args)
        if dev is None or dev == torch.cuda.current_device():
            return self.fn(*args)
        with torch.cuda.device(dev):
            return self.fn(*args)


class _Lib:
    def __init__(self, handle):
        self._handle = handle

    def __getattr__(self, name):
        fn = _Fn(getattr(self._handle, name), name)
        setattr(self, name, fn)
        return fn


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MpbError(f'{LIB_PATH} is missing: run ./build.sh (or __graft_entry__.build()). '
                           'There is no CPU fallback for the hot path.')
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        for which, struct in enumerate((RobotDesc, FieldDesc, GPDesc, ExtraCostDesc, NoiseDesc)):
            if handle.mpb_sizeof_desc(which) != C.sizeof(struct):
                raise MpbError(f'{struct.__name__}: ctypes layout ({C.sizeof(struct)} B) differs from the library '
                               f'({handle.mpb_sizeof_desc(which)} B); rebuild with ./build.sh')
        _lib = _Lib(handle)
    return _lib


_inited = set()


def init_device(device):
    """mpb_init() once per device (scheduler slots; must not happen lazily inside a CUDA-graph capture)."""
    idx = torch.device(device).index
    if idx is None:
        idx = torch.cuda.current_device()
    if idx in _inited:
        return
    with torch.cuda.device(idx):
        check(lib().mpb_init())
    _inited.add(idx)


def check(rc):
    if rc != 0:
        raise MpbError(f'libmpb_b200 error {rc}: {lib().mpb_last_error().decode()}')


def ptr(t):
    """Device pointer of a contiguous fp32/int32/uint8 CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise MpbError('expected a CUDA tensor: the hot path has no CPU implementation')
    if not t.is_contiguous():
        raise MpbError('expected a contiguous tensor')
    p = DevPtr(t.data_ptr())
    p.mpb_device = t.device.index if t.device.index is not None else torch.cuda.current_device()
    return p


def stream_ptr():
    """The stream argument of a C-ABI call: the current stream of the device the call's tensors live on."""
    return _STREAM


class NoiseStream:
    """Host-side state of the in-kernel noise: a seed and a draw counter, plus where this process sits in the job
    (first global particle / sample it owns, global count).  ``next()`` returns the descriptor of one draw and
    advances the counter, so successive optimize() iterations see fresh noise; two processes built with the same seed
    and the job's global extents draw exactly the slices of ONE global stream they own."""

    def __init__(self, seed=None, p_offset=0, P_global=1, s_offset=0, offset=0):
        self.seed = int(torch.initial_seed() if seed is None else seed) & 0xFFFFFFFFFFFFFFFF
        self.offset = int(offset)
        self.p_offset, self.P_global, self.s_offset = int(p_offset), int(P_global), int(s_offset)

    def desc(self, offset=None):
        return NoiseDesc(seed=self.seed, offset=self.offset if offset is None else int(offset), s_offset=self.s_offset,
                         p_offset=self.p_offset, P_global=self.P_global)

    def next(self):
        d = self.desc()
        self.offset += 1
        return d


def philox_normal(noise_desc, layout, shape, device, dof=None):
    """The normals a kernel called with ``noise_desc`` consumes, in the call's local layout (replay / debug; also the
    generator in front of the samplers without a fused variant).  NOISE_SPMD takes the dof count (``dof``)."""
    out = torch.empty(*shape, device=device, dtype=torch.float32)
    n = list(shape) + [1] * (4 - len(shape))
    if layout == NOISE_SPMD:
        n[3] = int(dof)
    check(lib().mpb_philox_normal(C.byref(noise_desc), layout, ptr(out), n[0], n[1], n[2], n[3], stream_ptr()))
    return out


def require_f32(*tensors):
    for t in tensors:
        if t is not None and t.dtype != torch.float32:
            raise MpbError(f'expected float32, got {t.dtype}')
