"""ctypes binding of libhande_b200.so (include/hande_b200.h): the only way this package computes anything.

There is no CPU fallback: if the CUDA library is missing or no GPU is visible, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

# values of the reference's excit_gen enumerators (src/qmc_data.f90:31-69)
EXCIT_GEN = {"renorm": 0, "renorm_spin": 1, "no_renorm": 2, "no_renorm_spin": 3, "power_pitzer": 4, "power_pitzer_occ": 5, "power_pitzer_occ_ij": 6, "power_pitzer_orderN": 7, "cauchy_schwarz_occ": 8,
             "cauchy_schwarz_occ_ij": 9, "heat_bath": 10, "heat_bath_uniform": 11, "heat_bath_single": 12}


class Config(C.Structure):
    _fields_ = [
        ("device", C.c_int32), ("nbasis", C.c_int32), ("nel", C.c_int32), ("excit_gen", C.c_int32),
        ("pattempt_single", C.c_double), ("pattempt_double", C.c_double),
        ("real_amplitudes", C.c_int32), ("spawn_cutoff", C.c_double),
        ("initiator_approx", C.c_int32), ("initiator_pop", C.c_double),
        ("trunc_level", C.c_int32),
        ("walker_length", C.c_int64), ("spawned_walker_length", C.c_int64),
        ("rng_seed", C.c_uint32), ("nprocs", C.c_int32), ("iproc", C.c_int32), ("nslots", C.c_int32),
        ("hash_seed", C.c_int32),
    ]


class SystemReadIn(C.Structure):
    _fields_ = [
        ("nbasis", C.c_int32), ("nel", C.c_int32), ("uhf", C.c_int32),
        ("nsym_tot", C.c_int32), ("sym0", C.c_int32), ("sym_max", C.c_int32), ("pg_mask", C.c_int32),
        ("Lz_mask", C.c_int32), ("Lz_offset", C.c_int32), ("gamma_sym", C.c_int32),
        ("nvirt", C.c_int32), ("nvirt_alpha", C.c_int32), ("nvirt_beta", C.c_int32), ("max_nbss", C.c_int32),
        ("Ecore", C.c_double),
        ("bf_sym", C.c_void_p), ("bf_ms", C.c_void_p), ("bf_spatial", C.c_void_p),
        ("nbasis_sym_spin", C.c_void_p), ("sym_spin_basis_fns", C.c_void_p),
        ("one_body", C.c_void_p), ("two_body", C.c_void_p * 4), ("nintgrls", C.c_int64),
    ]


class SystemUeg(C.Structure):
    _fields_ = [
        ("nbasis", C.c_int32), ("nel", C.c_int32), ("box_length", C.c_double),
        ("kvec", C.c_void_p), ("sp_eigv", C.c_void_p),
        ("kmax", C.c_int32), ("offset", C.c_int32), ("offset_inds", C.c_int32 * 3),
        ("lookup", C.c_void_p), ("n_lookup", C.c_int64),
        ("tern_kmax", C.c_int32), ("ternary_conserve", C.c_void_p),
    ]


class IterIn(C.Structure):
    _fields_ = [("tau", C.c_double), ("shift", C.c_double), ("proj_energy_old", C.c_double),
                ("first_cycle", C.c_uint32)]


class IterOut(C.Structure):
    _fields_ = [
        ("proj_energy", C.c_double), ("D0_population", C.c_double), ("nparticles", C.c_double),
        ("nstates", C.c_int64), ("nspawn_events", C.c_int64), ("ndeath", C.c_int64), ("nattempts", C.c_int64),
        ("rspawn", C.c_double), ("nattempts_spawn", C.c_int64), ("spawn_error", C.c_int32), ("psip_error", C.c_int32),
        ("walker_iterations", C.c_double),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class CcmcOut(C.Structure):
    _fields_ = [
        ("proj_energy", C.c_double), ("D0_population", C.c_double), ("D0_normalisation", C.c_double),
        ("tot_abs_real_pop", C.c_double), ("nattempts", C.c_int64), ("nattempts_spawn", C.c_int64),
        ("nspawn_events", C.c_int64), ("ndeath", C.c_int64), ("ndeath_nc", C.c_int64), ("spawn_error", C.c_int32),
        ("psip_error", C.c_int32),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_LIB = None


def _point_at_bundled_nccl():
    """Tell the engine where torch's bundled libnccl.so.2 lives (used only if no NCCL is loaded yet when
    hb200_comm_init runs), so that one process never ends up with two different NCCL builds."""
    if os.environ.get("HB200_NCCL_LIB"):
        return
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        if spec and spec.submodule_search_locations:
            cand = os.path.join(list(spec.submodule_search_locations)[0], "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["HB200_NCCL_LIB"] = cand
    except Exception:
        pass


def load_library():
    """Load (building if stale) the CUDA engine.  Raises if it cannot be built/loaded - never falls back."""
    global _LIB
    if _LIB is not None:
        return _LIB
    _point_at_bundled_nccl()
    path = _build.LIB
    if _build.needs_build():
        if os.path.exists("/usr/local/cuda/bin/nvcc"):
            _build.build()
        elif not os.path.exists(path):
            raise RuntimeError("libhande_b200.so is missing and nvcc is unavailable: the engine has no CPU fallback")
    L = C.CDLL(path)
    L.hb200_last_error.restype = C.c_char_p
    L.hb200_create.restype = C.c_void_p
    L.hb200_create.argtypes = [C.POINTER(Config)]
    L.hb200_destroy.argtypes = [C.c_void_p]
    L.hb200_set_system_read_in.argtypes = [C.c_void_p, C.POINTER(SystemReadIn)]
    L.hb200_set_system_ueg.argtypes = [C.c_void_p, C.POINTER(SystemUeg)]
    L.hb200_build_heat_bath.argtypes = [C.c_void_p]
    L.hb200_set_excit_tables.argtypes = [C.c_void_p, C.c_void_p]
    L.hb200_download_heat_bath.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64]
    L.hb200_set_reference.argtypes = [C.c_void_p, C.c_void_p, C.c_double]
    L.hb200_set_proc_map.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    L.hb200_upload_psips.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
    L.hb200_upload_psips_begin.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
    L.hb200_upload_psips_commit.argtypes = [C.c_void_p]
    L.hb200_download_psips.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    L.hb200_nstates.restype = C.c_int64
    L.hb200_nstates.argtypes = [C.c_void_p]
    L.hb200_iterate.argtypes = [C.c_void_p, C.c_int32, C.POINTER(IterIn), C.POINTER(IterOut)]
    L.hb200_spawn_death.argtypes = [C.c_void_p, C.POINTER(IterIn), C.c_uint32, C.POINTER(IterOut)]
    L.hb200_ccmc_spawn.argtypes = [C.c_void_p, C.POINTER(IterIn), C.c_uint32, C.c_int32, C.POINTER(CcmcOut)]
    L.hb200_ccmc_iterate.argtypes = [C.c_void_p, C.c_int32, C.POINTER(IterIn), C.c_int32, C.POINTER(IterOut)]
    L.hb200_ccmc_set_hash_shift.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
    L.hb200_ccmc_set_full_nc.argtypes = [C.c_void_p, C.c_int32]
    L.hb200_set_pattempt.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int32]
    L.hb200_get_ps_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    L.hb200_build_power_pitzer_orderN.argtypes = [C.c_void_p, C.c_double]
    L.hb200_build_power_pitzer.argtypes = [C.c_void_p, C.c_double]
    L.hb200_set_quasi_newton.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double]
    L.hb200_set_pattempt_parallel.argtypes = [C.c_void_p, C.c_double]
    L.hb200_get_pattempt_parallel.argtypes = [C.c_void_p]
    L.hb200_get_pattempt_parallel.restype = C.c_double
    L.hb200_comm_spawn.argtypes = [C.c_void_p]
    L.hb200_annihilate_spawn.argtypes = [C.c_void_p]
    L.hb200_annihilate_main.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(IterOut)]
    L.hb200_download_spawn.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    L.hb200_upload_spawn.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    L.hb200_spawn_counts.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    L.hb200_sc0_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    L.hb200_gen_excit_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_uint32,
                                        C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
    L.hb200_gen_excit_batch_rn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_double,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.hb200_get_unique_id.argtypes = [C.c_void_p]
    L.hb200_comm_init.argtypes = [C.c_void_p, C.c_void_p]
    L.hb200_p2p_export.argtypes = [C.c_void_p, C.c_void_p]
    L.hb200_set_propagator_weight.argtypes = [C.c_void_p, C.c_double]
    L.hb200_slot_populations.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    L.hb200_redistribute_particles.argtypes = [C.c_void_p, C.c_void_p]
    L.hb200_p2p_import.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    L.hb200_set_host_barrier.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.hb200_p2p_enable.argtypes = [C.c_void_p, C.c_int32]
    L.hb200_last_timing.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.hb200_set_determ_space.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.hb200_determ_hamil.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.hb200_determ_hamil.restype = C.c_int64
    L.hb200_determ_vector.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
    L.hb200_determ_project.argtypes = [C.c_void_p, C.POINTER(IterIn), C.c_uint32, C.c_void_p]
    _LIB = L
    return L


ABI_SYMBOLS = [
    "hb200_last_error", "hb200_create", "hb200_destroy", "hb200_set_system_read_in", "hb200_set_system_ueg",
    "hb200_build_heat_bath", "hb200_set_excit_tables", "hb200_gen_excit_batch_rn",
    "hb200_download_heat_bath", "hb200_set_reference", "hb200_set_proc_map", "hb200_upload_psips",
    "hb200_upload_psips_begin", "hb200_upload_psips_commit",
    "hb200_download_psips", "hb200_nstates", "hb200_iterate", "hb200_spawn_death", "hb200_comm_spawn",
    "hb200_ccmc_spawn", "hb200_ccmc_iterate", "hb200_ccmc_set_hash_shift", "hb200_ccmc_set_full_nc", "hb200_set_pattempt", "hb200_get_ps_stats", "hb200_set_pattempt_parallel", "hb200_get_pattempt_parallel", "hb200_build_power_pitzer_orderN", "hb200_build_power_pitzer", "hb200_set_quasi_newton", "hb200_annihilate_spawn", "hb200_annihilate_main", "hb200_download_spawn", "hb200_upload_spawn", "hb200_spawn_counts",
    "hb200_sc0_batch", "hb200_gen_excit_batch", "hb200_get_unique_id", "hb200_comm_init", "hb200_last_timing",
    "hb200_p2p_export", "hb200_p2p_import", "hb200_set_host_barrier", "hb200_p2p_enable", "hb200_slot_populations", "hb200_set_propagator_weight",
    "hb200_redistribute_particles", "hb200_set_determ_space", "hb200_determ_hamil", "hb200_determ_vector",
    "hb200_determ_project",
]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class EngineError(RuntimeError):
    pass


# field order of hb200_heat_bath_tables (include/hande_b200.h)
HEAT_BATH_TABLE_KEYS = ("i_weights", "ij_weights", "ija_weights", "ija_aliasU", "ija_aliasK", "ija_weights_tot",
                        "ijab_weights", "ijab_aliasU", "ijab_aliasK", "ijab_weights_tot")


class Engine:
    """One GPU's share of the walker population (one MPI rank of the reference)."""

    def __init__(self, sys, *, excit_gen="renorm", pattempt_single, pattempt_double, real_amplitudes=False,
                 spawn_cutoff=0.01, initiator_approx=False, initiator_pop=3.0, trunc_level=-1, walker_length=1 << 20,
                 spawned_walker_length=1 << 18, seed=7, nprocs=1, iproc=0, nslots=1, device=0, hash_seed=7,
                 pattempt_parallel=-1.0, power_pitzer_min_weight=0.01, heat_bath_tables=None):
        """heat_bath_tables: the ten arrays of excit_gen_heat_bath_t as a host built them (dict with the keys of
        HEAT_BATH_TABLE_KEYS, the reference's column-major order) - handed over with hb200_set_excit_tables instead of
        being built on the device"""
        self.L = load_library()
        self.power_pitzer_min_weight = power_pitzer_min_weight
        self.sys = sys
        self.W = sys.W
        self.E = sys.W + 2
        self.real_factor = (1 << 31) if real_amplitudes else 1
        self.nprocs, self.iproc, self.nslots = nprocs, iproc, nslots
        eg = EXCIT_GEN[excit_gen] if isinstance(excit_gen, str) else int(excit_gen)
        self.cfg = Config(device=device, nbasis=sys.nbasis, nel=sys.nel, excit_gen=eg,
                          pattempt_single=pattempt_single, pattempt_double=pattempt_double,
                          real_amplitudes=int(real_amplitudes), spawn_cutoff=spawn_cutoff,
                          initiator_approx=int(initiator_approx), initiator_pop=initiator_pop, trunc_level=trunc_level,
                          walker_length=walker_length, spawned_walker_length=spawned_walker_length, rng_seed=seed,
                          nprocs=nprocs, iproc=iproc, nslots=nslots, hash_seed=hash_seed)
        h = self.L.hb200_create(C.byref(self.cfg))
        if not h:
            raise EngineError(self.L.hb200_last_error().decode())
        self.h = C.c_void_p(h)
        self._set_system(sys)
        if eg in (EXCIT_GEN["heat_bath"], EXCIT_GEN["heat_bath_uniform"], EXCIT_GEN["heat_bath_single"],
                  EXCIT_GEN["power_pitzer_occ_ij"],
                  EXCIT_GEN["cauchy_schwarz_occ_ij"]):   # the _occ_ij weights are the heat-bath i/ij tables
            if heat_bath_tables is None:
                self._chk(self.L.hb200_build_heat_bath(self.h))
            else:
                keep = [np.ascontiguousarray(heat_bath_tables[k], dtype=(np.int32 if k.endswith("aliasK") else np.float64))
                        for k in HEAT_BATH_TABLE_KEYS]
                ptrs = (C.c_void_p * len(keep))(*[a.ctypes.data for a in keep])
                self._chk(self.L.hb200_set_excit_tables(self.h, C.cast(ptrs, C.c_void_p)))
        if eg in (EXCIT_GEN["renorm_spin"], EXCIT_GEN["no_renorm_spin"]):
            self.pattempt_parallel = self.set_pattempt_parallel(pattempt_parallel)

    def close(self):
        if getattr(self, "h", None):
            self.L.hb200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc != 0:
            raise EngineError(self.L.hb200_last_error().decode())

    def _set_system(self, s):
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)  # noqa: E731
        if getattr(s, "kind", "read_in") == "ueg":
            keep = [i32(s.kvec), np.ascontiguousarray(s.sp_eigv, dtype=np.float64), i32(s.lookup),
                    np.ascontiguousarray(s.ternary_conserve, dtype=np.uint64)]
            su = SystemUeg(nbasis=s.nbasis, nel=s.nel, box_length=s.L, kvec=keep[0].ctypes.data,
                           sp_eigv=keep[1].ctypes.data, kmax=s.kmax, offset=s.offset,
                           offset_inds=(C.c_int32 * 3)(*[int(x) for x in s.offset_inds]), lookup=keep[2].ctypes.data,
                           n_lookup=len(keep[2]) - 1, tern_kmax=s.tern_kmax, ternary_conserve=keep[3].ctypes.data)
            self._chk(self.L.hb200_set_system_ueg(self.h, C.byref(su)))
            return
        keep = [i32(s.sym), i32(s.ms), i32(s.spatial), i32(s.nbasis_sym_spin), i32(s.sym_spin_basis_fns),
                np.ascontiguousarray(s.h1, dtype=np.float64)]
        v2 = [np.ascontiguousarray(v, dtype=np.float64) for v in s.v2]
        tb = (C.c_void_p * 4)(*[v.ctypes.data for v in v2] + [None] * (4 - len(v2)))
        si = SystemReadIn(nbasis=s.nbasis, nel=s.nel, uhf=int(s.uhf), nsym_tot=s.nsym_tot, sym0=s.sym0,
                          sym_max=s.sym_max, pg_mask=s.pg_mask, Lz_mask=s.Lz_mask, Lz_offset=s.Lz_offset,
                          gamma_sym=s.gamma_sym, nvirt=s.nvirt, nvirt_alpha=s.nvirt_alpha, nvirt_beta=s.nvirt_beta,
                          max_nbss=s.max_nbss, Ecore=s.Ecore, bf_sym=keep[0].ctypes.data, bf_ms=keep[1].ctypes.data,
                          bf_spatial=keep[2].ctypes.data, nbasis_sym_spin=keep[3].ctypes.data,
                          sym_spin_basis_fns=keep[4].ctypes.data, one_body=keep[5].ctypes.data, two_body=tb,
                          nintgrls=s.nintgrls)
        self._chk(self.L.hb200_set_system_read_in(self.h, C.byref(si)))

    # ---- reference / lists
    def set_reference(self, f0, H00):
        f0 = np.ascontiguousarray(f0, dtype=np.uint64)
        self._chk(self.L.hb200_set_reference(self.h, _p(f0), float(H00)))
        if self.cfg.excit_gen == EXCIT_GEN["power_pitzer_orderN"]:   # reference-mapped tables (src/qmc.F90:1009-1018)
            self._chk(self.L.hb200_build_power_pitzer_orderN(self.h, float(self.power_pitzer_min_weight)))
        if self.cfg.excit_gen == EXCIT_GEN["power_pitzer"]:          # src/qmc.F90:994-1007
            self._chk(self.L.hb200_build_power_pitzer(self.h, float(self.power_pitzer_min_weight)))

    def set_proc_map(self, pmap):
        pmap = np.ascontiguousarray(pmap, dtype=np.int32)
        self._chk(self.L.hb200_set_proc_map(self.h, _p(pmap), len(pmap)))

    def upload_psips(self, states, pops, dat):
        states = np.ascontiguousarray(states, dtype=np.uint64).reshape(-1, self.W)
        pops = np.ascontiguousarray(pops, dtype=np.int64)
        dat = np.ascontiguousarray(dat, dtype=np.float64)
        assert len(states) == len(pops) == len(dat)
        self._chk(self.L.hb200_upload_psips(self.h, _p(states), _p(pops), _p(dat), len(pops)))

    def upload_psips_ptr(self, states_ptr, pops_ptr, dat_ptr, n):
        """hb200_upload_psips from caller-owned (e.g. pinned) host buffers given as integer addresses."""
        self._chk(self.L.hb200_upload_psips(self.h, C.c_void_p(states_ptr), C.c_void_p(pops_ptr), C.c_void_p(dat_ptr), n))

    def upload_psips_begin_ptr(self, states_ptr, pops_ptr, dat_ptr, n):
        """Start the asynchronous upload of the next list (pinned host buffers) while the current one propagates."""
        self._chk(self.L.hb200_upload_psips_begin(self.h, C.c_void_p(states_ptr), C.c_void_p(pops_ptr),
                                                  C.c_void_p(dat_ptr), n))

    def upload_psips_commit(self):
        self._chk(self.L.hb200_upload_psips_commit(self.h))

    def download_psips_ptr(self, states_ptr, pops_ptr, dat_ptr, capacity):
        nn = C.c_int64(0)
        self._chk(self.L.hb200_download_psips(self.h, C.c_void_p(states_ptr), C.c_void_p(pops_ptr),
                                              C.c_void_p(dat_ptr), capacity, C.byref(nn)))
        return nn.value

    @property
    def nstates(self):
        return int(self.L.hb200_nstates(self.h))

    def download_psips(self):
        n = self.nstates
        states = np.zeros((n, self.W), dtype=np.uint64)
        pops = np.zeros(n, dtype=np.int64)
        dat = np.zeros(n)
        nn = C.c_int64(0)
        self._chk(self.L.hb200_download_psips(self.h, _p(states), _p(pops), _p(dat), n, C.byref(nn)))
        return states, pops, dat

    # ---- propagation
    def iterate(self, ncycles, tau, shift, proj_energy_old, first_cycle):
        i = IterIn(tau=tau, shift=shift, proj_energy_old=proj_energy_old, first_cycle=first_cycle)
        o = IterOut()
        self._chk(self.L.hb200_iterate(self.h, ncycles, C.byref(i), C.byref(o)))
        return o.as_dict()

    def spawn_death(self, tau, shift, proj_energy_old, cycle):
        i = IterIn(tau=tau, shift=shift, proj_energy_old=proj_energy_old, first_cycle=cycle)
        o = IterOut()
        self._chk(self.L.hb200_spawn_death(self.h, C.byref(i), cycle, C.byref(o)))
        return o.as_dict()

    # ---- CCMC
    def ccmc_spawn(self, tau, shift, proj_energy_old, cycle, ex_level):
        i = IterIn(tau=tau, shift=shift, proj_energy_old=proj_energy_old, first_cycle=cycle)
        o = CcmcOut()
        self._chk(self.L.hb200_ccmc_spawn(self.h, C.byref(i), cycle, ex_level, C.byref(o)))
        return o.as_dict()

    def ccmc_iterate(self, ncycles, tau, shift, proj_energy_old, first_cycle, ex_level):
        i = IterIn(tau=tau, shift=shift, proj_energy_old=proj_energy_old, first_cycle=first_cycle)
        o = IterOut()
        self._chk(self.L.hb200_ccmc_iterate(self.h, ncycles, C.byref(i), ex_level, C.byref(o)))
        return o.as_dict()

    def ccmc_set_full_nc(self, full_nc=True):
        self._chk(self.L.hb200_ccmc_set_full_nc(self.h, int(bool(full_nc))))

    def set_pattempt(self, pattempt_single, pattempt_double, accumulate=False):
        """excit_gen_data%pattempt_single/double; accumulate: collect the pattempt_update statistics on the device"""
        self._chk(self.L.hb200_set_pattempt(self.h, float(pattempt_single), float(pattempt_double), int(bool(accumulate))))

    def set_quasi_newton(self, sp_fock, ref_fock_sum, threshold, value, pop_control):
        """qmc = { quasi_newton = true }: propagator%sp_fock (1-based, entry 0 unused) and the resolved scalars"""
        sp = np.ascontiguousarray(sp_fock, dtype=np.float64)
        assert len(sp) == self.sys.nbasis + 1
        self._chk(self.L.hb200_set_quasi_newton(self.h, _p(sp), float(ref_fock_sum), float(threshold), float(value),
                                                float(pop_control)))

    # ---- semi-stochastic projection (src/semi_stoch.F90)
    def set_determ_space(self, dets, sizes):
        """init_semi_stoch_t with the host's space: dets = determ%dets (tot x W, rank by rank, each rank's part in list
        order), sizes = determ%sizes; all sizes zero switches the projection off"""
        sizes = np.ascontiguousarray(sizes, dtype=np.int32)
        assert len(sizes) == self.nprocs
        dets = np.ascontiguousarray(dets, dtype=np.uint64).reshape(-1, self.W)
        assert len(dets) == int(sizes.sum())
        self.determ_sizes = sizes.copy()
        self._chk(self.L.hb200_set_determ_space(self.h, _p(dets) if len(dets) else None, _p(sizes)))

    def determ_hamil(self):
        """determ%hamil of this rank by column: (col_ptr, row, val)"""
        nnz = self.L.hb200_determ_hamil(self.h, None, None, None)
        nloc = int(self.determ_sizes[self.iproc])
        cp = np.zeros(nloc + 1, dtype=np.int64)
        row = np.zeros(max(nnz, 1), dtype=np.int32)
        val = np.zeros(max(nnz, 1))
        if self.L.hb200_determ_hamil(self.h, _p(cp), _p(row), _p(val)) < 0:
            raise EngineError("determ_hamil")
        return cp, row[:nnz], val[:nnz]

    def determ_vector(self, which=0):
        """determ%vector of this rank: which=0 the deterministic amplitudes now, which=1 the last projection's result"""
        v = np.zeros(int(self.determ_sizes[self.iproc]))
        self._chk(self.L.hb200_determ_vector(self.h, int(which), _p(v) if len(v) else None))
        return v

    def determ_project(self, tau, shift, proj_energy_old, cycle, full_vector=None):
        """staged determ_projection + deterministic_annihilation (between spawn_death and annihilate_main)"""
        i = IterIn(tau=tau, shift=shift, proj_energy_old=proj_energy_old, first_cycle=cycle)
        fv = None if full_vector is None else np.ascontiguousarray(full_vector, dtype=np.float64)
        self._chk(self.L.hb200_determ_project(self.h, C.byref(i), cycle, None if fv is None else _p(fv)))

    def set_propagator_weight(self, weight):
        """wall-Chebyshev: weight 1/(S_i - E_0) of the sub-cycle the next iterate()/spawn_death() calls run"""
        self._chk(self.L.hb200_set_propagator_weight(self.h, float(weight)))

    def set_pattempt_parallel(self, pattempt_parallel=-1.0):
        """qmc_in%pattempt_parallel (renorm_spin / no_renorm_spin); negative: find_parallel_spin_prob_mol on the device"""
        self._chk(self.L.hb200_set_pattempt_parallel(self.h, float(pattempt_parallel)))
        return float(self.L.hb200_get_pattempt_parallel(self.h))

    def get_ps_stats(self, reset=True):
        """this rank's (h_pgen_singles_sum, excit_gen_singles, h_pgen_doubles_sum, excit_gen_doubles) since the last reset"""
        out = np.zeros(4)
        self._chk(self.L.hb200_get_ps_stats(self.h, _p(out), int(bool(reset))))
        return out

    def ccmc_set_hash_shift(self, hash_shift, move_freq=5):
        self._chk(self.L.hb200_ccmc_set_hash_shift(self.h, int(hash_shift), int(move_freq)))

    def slot_populations(self):
        out = np.zeros(self.nprocs * self.nslots)
        self._chk(self.L.hb200_slot_populations(self.h, _p(out), len(out)))
        return out

    def redistribute_particles(self):
        """redistribute_particles after set_proc_map; continue with comm_spawn / annihilate_spawn / annihilate_main
        (or redistribute() for all four).  Returns the population that left this rank."""
        ns = C.c_double(0.0)
        self._chk(self.L.hb200_redistribute_particles(self.h, C.byref(ns)))
        return ns.value

    def redistribute(self, cycle):
        """redistribute_load_balancing_dets (src/qmc_common.F90:1332-1390) through the NCCL exchange."""
        self.redistribute_particles()
        self.comm_spawn()
        self.annihilate_spawn()
        return self.annihilate_main(cycle)

    def comm_spawn(self):
        self._chk(self.L.hb200_comm_spawn(self.h))

    def annihilate_spawn(self):
        self._chk(self.L.hb200_annihilate_spawn(self.h))

    def annihilate_main(self, cycle):
        o = IterOut()
        self._chk(self.L.hb200_annihilate_main(self.h, cycle, C.byref(o)))
        return o.as_dict()

    def download_spawn(self):
        cap = int(self.cfg.spawned_walker_length)
        buf = np.zeros((cap, self.E), dtype=np.int64)
        n = C.c_int64(0)
        self._chk(self.L.hb200_download_spawn(self.h, _p(buf), cap, C.byref(n)))
        return buf[: n.value].copy()

    def spawn_counts(self):
        """elements per destination block of the spawn list after the spawning stage (spawn%head)"""
        c = np.zeros(int(self.cfg.nprocs), dtype=np.int64)
        self._chk(self.L.hb200_spawn_counts(self.h, _p(c), len(c)))
        return c

    def upload_spawn(self, sdata):
        sdata = np.ascontiguousarray(sdata, dtype=np.int64).reshape(-1, self.E)
        self._chk(self.L.hb200_upload_spawn(self.h, _p(sdata), len(sdata)))

    # ---- pure-function batches
    def sc0_batch(self, states):
        states = np.ascontiguousarray(states, dtype=np.uint64).reshape(-1, self.W)
        out = np.zeros(len(states))
        self._chk(self.L.hb200_sc0_batch(self.h, _p(states), len(states), _p(out)))
        return out

    def gen_excit_batch(self, states, pops, attempts, cycle, tau):
        states = np.ascontiguousarray(states, dtype=np.uint64).reshape(-1, self.W)
        pops = np.ascontiguousarray(pops, dtype=np.int64)
        attempts = np.ascontiguousarray(attempts, dtype=np.uint32)
        n = len(pops)
        io = np.zeros((n, 8), dtype=np.int32)
        do = np.zeros((n, 2))
        ns = np.zeros(n, dtype=np.int64)
        self._chk(self.L.hb200_gen_excit_batch(self.h, _p(states), _p(pops), _p(attempts), n, cycle, tau, _p(io), _p(do),
                                               _p(ns)))
        return io, do, ns

    def gen_excit_batch_rn(self, states, pops, rn, tau):
        """gen_excit + attempt_to_spawn on injected random numbers rn[n][nrn]; returns (iout, dout, nspawn, nused)"""
        states = np.ascontiguousarray(states, dtype=np.uint64).reshape(-1, self.W)
        pops = np.ascontiguousarray(pops, dtype=np.int64)
        rn = np.ascontiguousarray(rn, dtype=np.float64)
        n = len(pops)
        io = np.zeros((n, 8), dtype=np.int32)
        do = np.zeros((n, 2))
        ns = np.zeros(n, dtype=np.int64)
        nu = np.zeros(n, dtype=np.int32)
        self._chk(self.L.hb200_gen_excit_batch_rn(self.h, _p(states), _p(pops), _p(rn), rn.shape[1], n, tau, _p(io), _p(do),
                                                  _p(ns), _p(nu)))
        return io, do, ns, nu

    def heat_bath_table(self, which, n):
        out = np.zeros(n, dtype=np.int32 if which >= 8 else np.float64)
        self._chk(self.L.hb200_download_heat_bath(self.h, which, _p(out), n))
        return out

    # ---- multi-GPU
    @staticmethod
    def get_unique_id():
        L = load_library()
        buf = np.zeros(128, dtype=np.uint8)
        if L.hb200_get_unique_id(_p(buf)) != 0:
            raise EngineError(L.hb200_last_error().decode())
        return buf

    def comm_init(self, uid):
        uid = np.ascontiguousarray(uid, dtype=np.uint8)
        self._chk(self.L.hb200_comm_init(self.h, _p(uid)))

    def comm_setup(self, comm, p2p=None, nccl=True):
        """Everything a multi-rank engine needs before hb200_iterate: the NCCL communicator (unique id broadcast from
        rank 0) and, unless disabled (p2p=False or HB200_NO_P2P=1), the peer-to-peer receive buffers (CUDA IPC
        handles all-gathered through `comm`).  `comm` offers broadcast_bytes / allgather_bytes (TorchDist).
        nccl=False: no NCCL communicator; exchanges end with the host's barrier (`comm.barrier`) - the way a plain-MPI
        host, or several ranks sharing one GPU, run the peer-to-peer exchange."""
        if nccl:
            uid = self.get_unique_id() if comm.rank == 0 else np.zeros(128, dtype=np.uint8)
            self.comm_init(comm.broadcast_bytes(uid, src=0))
        else:
            self._barrier_cb = C.CFUNCTYPE(None, C.c_void_p)(lambda _arg: comm.barrier())   # kept alive with the engine
            self._chk(self.L.hb200_set_host_barrier(self.h, C.cast(self._barrier_cb, C.c_void_p), None))
            p2p = True
        if p2p is None:
            p2p = os.environ.get("HB200_NO_P2P", "0") in ("", "0")
        self.p2p = False
        if p2p:
            h = np.zeros(64, dtype=np.uint8)
            self._chk(self.L.hb200_p2p_export(self.h, _p(h)))
            allh = np.ascontiguousarray(comm.allgather_bytes(h), dtype=np.uint8)
            ok = self.L.hb200_p2p_import(self.h, _p(allh), comm.size) == 0
            # every rank must run the same exchange: peer-to-peer only if the import succeeded everywhere
            nok = int(round(float(comm.allreduce_sum(np.array([1.0 if ok else 0.0]))[0])))
            if nok == comm.size:
                self.p2p = True
            else:
                if not nccl:
                    raise EngineError("peer-to-peer import failed and no NCCL communicator to fall back to: " +
                                      self.L.hb200_last_error().decode())
                if ok:
                    self._chk(self.L.hb200_p2p_enable(self.h, 0))

    def last_timing(self):
        ms = np.zeros(8)
        cnt = np.zeros(4, dtype=np.int64)
        self.L.hb200_last_timing(self.h, _p(ms), _p(cnt))
        return {"spawn_ms": ms[0], "comm_ms": ms[1], "sort_ms": ms[2], "annihilate_ms": ms[3], "total_ms": ms[4], "spawn_kernel_ms": ms[5],
                "spawn_launches": int(cnt[0]), "launches": int(cnt[1])}
