"""ORACLE (test infrastructure, NOT product code): ctypes driver for oracle/liboracle.so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
LIB_WIDE = os.path.join(HERE, "liboracle_wide.so")     # MAXW = 32 (bit strings of up to 2048 spin-orbitals)
REF_LIB = os.path.join(HERE, "_ref", "libhande_ref_c.so")
HUGE = 2**31 - 1

EXCIT_GEN = {"renorm": 0, "renorm_spin": 1, "no_renorm": 2, "no_renorm_spin": 3, "power_pitzer": 4, "power_pitzer_occ": 5, "power_pitzer_occ_ij": 6, "power_pitzer_orderN": 7, "cauchy_schwarz_occ": 8,
             "cauchy_schwarz_occ_ij": 9, "heat_bath": 10,
             "heat_bath_uniform": 11, "heat_bath_single": 12}


class QmcIn(C.Structure):
    _fields_ = [
        ("tau", C.c_double), ("seed", C.c_int), ("D0_population", C.c_double),
        ("ncycles", C.c_int), ("nreport", C.c_int),
        ("target_particles", C.c_double), ("initial_shift", C.c_double), ("shift_damping", C.c_double),
        ("vary_shift_from", C.c_double), ("vary_shift_from_proje", C.c_int),
        ("initiator_approx", C.c_int), ("initiator_pop", C.c_double),
        ("real_amplitudes", C.c_int), ("spawn_cutoff", C.c_double),
        ("excit_gen", C.c_int), ("pattempt_single", C.c_double), ("pattempt_double", C.c_double),
        ("walker_length", C.c_int64), ("spawned_walker_length", C.c_int64),
        ("ex_level", C.c_int), ("nprocs", C.c_int), ("nslots", C.c_int), ("rng_kind", C.c_int),
        ("literal_event_int32", C.c_int),
    ]


def build(force=False):
    """Compile the oracle (and oracle/_ref when the reference tree is present)."""
    for target in (LIB, LIB_WIDE):
        if force or not os.path.exists(target) or any(
                os.path.getmtime(os.path.join(HERE, f)) > os.path.getmtime(target)
                for f in ("capi.cpp", "system.hpp", "rng.hpp", "excit_gen.hpp", "ueg.hpp", "fciqmc.hpp", "ccmc.hpp")):
            subprocess.check_call(["make", "-C", HERE, os.path.basename(target)], stdout=subprocess.DEVNULL)
    if not os.path.exists(REF_LIB) and os.path.isdir("/root/reference/lib/dSFMT-src-2.2.3"):
        subprocess.check_call(["make", "-C", HERE, "ref"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


_libs = {}


def lib(wide=False):
    if wide not in _libs:
        build()
        L = C.CDLL(LIB_WIDE if wide else LIB)
        L.orc_last_error.restype = C.c_char_p
        L.orc_create.restype = C.c_void_p
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_ecore.restype = C.c_double
        L.orc_ecore.argtypes = [C.c_void_p]
        for name in ("orc_sc0", "orc_proj_hmatel"):
            getattr(L, name).restype = C.c_double
            getattr(L, name).argtypes = [C.c_void_p, C.c_void_p]
        for name in ("orc_sc1", "orc_sc1_checked"):
            getattr(L, name).restype = C.c_double
            getattr(L, name).argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        for name in ("orc_sc2", "orc_sc2_checked"):
            getattr(L, name).restype = C.c_double
            getattr(L, name).argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_get_one_body.restype = C.c_double
        L.orc_get_one_body.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_get_two_body.restype = C.c_double
        L.orc_get_two_body.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_two_body_store.restype = C.POINTER(C.c_double)
        L.orc_two_body_store.argtypes = [C.c_void_p, C.c_int]
        L.orc_read_fcidump.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_sys_info.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_init_ueg.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double]
        L.orc_set_ref_det_list.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.orc_ueg_info.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_ueg_lookup.restype = C.POINTER(C.c_int)
        L.orc_ueg_lookup.argtypes = [C.c_void_p]
        L.orc_ueg_ternary.restype = C.POINTER(C.c_uint64)
        L.orc_ueg_ternary.argtypes = [C.c_void_p]
        L.orc_basis.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        L.orc_sym_tables.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_murmur_bit_string.restype = C.c_int32
        L.orc_murmur_bit_string.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        L.orc_murmur2.restype = C.c_uint32
        L.orc_murmur2.argtypes = [C.c_void_p, C.c_int, C.c_uint32]
        L.orc_ref_murmur2.restype = C.c_uint32
        L.orc_ref_murmur2.argtypes = [C.c_void_p, C.c_int, C.c_uint32]
        L.orc_owner.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_dsfmt_stream.argtypes = [C.c_int, C.c_int, C.c_void_p]
        L.orc_philox_stream.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_int, C.c_uint32,
                                        C.c_int, C.c_void_p]
        L.orc_set_qmc.argtypes = [C.c_void_p, C.POINTER(QmcIn)]
        L.orc_init.argtypes = [C.c_void_p]
        L.orc_run.argtypes = [C.c_void_p]
        L.orc_run_ccmc.argtypes = [C.c_void_p]
        L.orc_get_nattempts_rows.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_nrows.argtypes = [C.c_void_p]
        L.orc_get_rows.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_reference.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_nstates.restype = C.c_int64
        L.orc_nstates.argtypes = [C.c_void_p, C.c_int]
        L.orc_nparticles.restype = C.c_double
        L.orc_nparticles.argtypes = [C.c_void_p, C.c_int]
        L.orc_get_psips.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_set_psips.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_set_reference_det.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_iterate.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_double, C.c_double, C.c_double, C.c_void_p]
        L.orc_stage_spawn.argtypes = [C.c_void_p, C.c_uint32, C.c_double, C.c_double, C.c_double, C.c_void_p]
        L.orc_ccmc_set_full_nc.argtypes = [C.c_void_p, C.c_int]
        L.orc_ccmc_stage_spawn.argtypes = [C.c_void_p, C.c_uint32, C.c_double, C.c_double, C.c_double, C.c_void_p]
        L.orc_spawn_count.restype = C.c_int64
        L.orc_spawn_count.argtypes = [C.c_void_p, C.c_int]
        L.orc_get_spawn.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_stage_annihilate.argtypes = [C.c_void_p]
        L.orc_gen_excit_philox.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int64, C.c_double,
                                           C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_gen_excit_list.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_hb_nb.restype = C.c_int64
        L.orc_hb_nb.argtypes = [C.c_void_p]
        L.orc_hb_ptr_d.restype = C.POINTER(C.c_double)
        L.orc_hb_ptr_d.argtypes = [C.c_void_p, C.c_int]
        L.orc_hb_ptr_i.restype = C.POINTER(C.c_int)
        L.orc_hb_ptr_i.argtypes = [C.c_void_p, C.c_int]
        L.orc_cpu_baseline.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_void_p]
        _libs[wide] = L
    return _libs[wide]


def have_ref_lib():
    return os.path.exists(REF_LIB)


def use_ref_lib():
    if lib().orc_set_ref_lib(REF_LIB.encode()) != 0:
        raise RuntimeError(lib().orc_last_error().decode())


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    """One restated HANDE calculation (system + FCIQMC state)."""

    def __init__(self, wide=False):
        """wide=True: the 32-word build (bit strings of more than 256 spin-orbitals)"""
        self.L = lib(wide)
        self.h = C.c_void_p(self.L.orc_create())

    def __del__(self):
        try:
            self.L.orc_destroy(self.h)
        except Exception:
            pass

    def _chk(self, rc):
        if rc < 0:
            raise RuntimeError(self.L.orc_last_error().decode())
        return rc

    def init_ueg(self, nel, ms, rs, cutoff):
        """sys = ueg { electrons = nel, ms = ms, dim = 3, cutoff = cutoff, rs = rs }"""
        self._chk(self.L.orc_init_ueg(self.h, nel, ms, float(rs), float(cutoff)))
        return self._read_info()

    def set_ref_det(self, occ):
        """reference = { det = {...} } (1-based spin-orbital list); call before init()."""
        occ = np.ascontiguousarray(sorted(occ), dtype=np.int32)
        self.L.orc_set_ref_det_list(self.h, _p(occ), len(occ))

    def ueg_tables(self):
        kvec = np.zeros((self.nbasis, 3), dtype=np.int32)
        info = np.zeros(16, dtype=np.int64)
        dinfo = np.zeros(4)
        self.L.orc_ueg_info(self.h, _p(kvec), _p(info), _p(dinfo))
        lookup = np.ctypeslib.as_array(self.L.orc_ueg_lookup(self.h), shape=(int(info[5]),)).copy()
        tern = np.ctypeslib.as_array(self.L.orc_ueg_ternary(self.h), shape=(int(info[8]),)).copy()
        return {"kvec": kvec, "kmax": int(info[0]), "offset": int(info[1]), "offset_inds": info[2:5].astype(np.int32),
                "lookup": lookup, "tK": int(info[6]), "tD": int(info[7]), "ternary": tern, "L": dinfo[0],
                "rs": dinfo[1], "ecutoff": dinfo[2]}

    def read_fcidump(self, path, nel=0, ms=HUGE, sym=HUGE, cas=(-1, -1)):
        self._chk(self.L.orc_read_fcidump(self.h, str(path).encode(), nel, ms, sym, cas[0], cas[1]))
        return self._read_info()

    def _read_info(self):
        info = np.zeros(32, dtype=np.int64)
        self.L.orc_sys_info(self.h, _p(info))
        keys = ["nbasis", "nel", "W", "nsym_tot", "sym0", "sym_max", "nalpha", "nbeta", "uhf", "pg_mask", "Lz_mask",
                "Lz_offset", "gamma_sym", "max_nbss", "nvirt", "nvirt_alpha", "nvirt_beta", "symmetry", "int_err",
                "nchan", "nintgrls"]
        self.info = {k: int(v) for k, v in zip(keys, info)}
        self.nbasis, self.nel, self.W = self.info["nbasis"], self.info["nel"], self.info["W"]
        return self.info

    @property
    def ecore(self):
        return self.L.orc_ecore(self.h)

    def basis(self):
        nb = self.nbasis
        out = {k: np.zeros(nb, dtype=np.int32) for k in ("sym", "ms", "spatial", "sym_index", "sym_spin_index")}
        eig = np.zeros(nb)
        self.L.orc_basis(self.h, _p(out["sym"]), _p(out["ms"]), _p(out["spatial"]), _p(out["sym_index"]),
                         _p(out["sym_spin_index"]), _p(eig))
        out["sp_eigv"] = eig
        return out

    def sym_tables(self):
        ns, mx = self.info["nsym_tot"], self.info["max_nbss"]
        nbss = np.zeros(2 * ns, dtype=np.int32)
        ssbf = np.zeros(mx * 2 * ns, dtype=np.int32)
        self.L.orc_sym_tables(self.h, _p(nbss), _p(ssbf))
        return nbss, ssbf

    def two_body_store(self, chan=0):
        n = self.info["nintgrls"]
        return np.ctypeslib.as_array(self.L.orc_two_body_store(self.h, chan), shape=(n,)).copy()

    def one_body(self, i, j):
        return self.L.orc_get_one_body(self.h, i, j)

    def two_body(self, i, j, a, b):
        return self.L.orc_get_two_body(self.h, i, j, a, b)

    def det(self, occ):
        f = np.zeros(self.W, dtype=np.uint64)
        for o in occ:
            f[(o - 1) // 64] |= np.uint64(1) << np.uint64((o - 1) % 64)
        return f

    def sc0(self, f):
        f = np.ascontiguousarray(f, dtype=np.uint64)
        return self.L.orc_sc0(self.h, _p(f))

    def proj_hmatel(self, f):
        f = np.ascontiguousarray(f, dtype=np.uint64)
        return self.L.orc_proj_hmatel(self.h, _p(f))

    def get_hmatel(self, f1, f2):
        f1 = np.ascontiguousarray(f1, dtype=np.uint64)
        f2 = np.ascontiguousarray(f2, dtype=np.uint64)
        self.L.orc_get_hmatel.restype = C.c_double
        self.L.orc_get_hmatel.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        return self.L.orc_get_hmatel(self.h, _p(f1), _p(f2))

    def sc1(self, f, i, a, checked=False):
        f = np.ascontiguousarray(f, dtype=np.uint64)
        fn = self.L.orc_sc1_checked if checked else self.L.orc_sc1
        return fn(self.h, _p(f), i, a)

    def sc2(self, f, i, j, a, b, checked=False):
        f = np.ascontiguousarray(f, dtype=np.uint64)
        fn = self.L.orc_sc2_checked if checked else self.L.orc_sc2
        return fn(self.h, _p(f), i, j, a, b)

    def owner(self, f):
        f = np.ascontiguousarray(f, dtype=np.uint64)
        return self.L.orc_owner(self.h, _p(f))

    def murmur_bit_string(self, f, seed=7):
        f = np.ascontiguousarray(f, dtype=np.uint64)
        return self.L.orc_murmur_bit_string(self.h, _p(f), seed)

    def set_qmc(self, **kw):
        q = QmcIn(tau=0.001, seed=7, D0_population=10.0, ncycles=20, nreport=10, target_particles=1e7,
                  initial_shift=0.0, shift_damping=0.05, vary_shift_from=0.0, vary_shift_from_proje=0,
                  initiator_approx=0, initiator_pop=3.0, real_amplitudes=0, spawn_cutoff=0.01, excit_gen=EXCIT_GEN["renorm"],
                  pattempt_single=-1.0, pattempt_double=-1.0, walker_length=1 << 20,
                  spawned_walker_length=1 << 18, ex_level=-1, nprocs=1, nslots=1, rng_kind=0,
                  literal_event_int32=1)
        for k, v in kw.items():
            if k == "excit_gen" and isinstance(v, str):
                v = EXCIT_GEN[v]
            if not hasattr(q, k):
                raise KeyError(k)
            setattr(q, k, v)
        self.qmc = q
        self.L.orc_set_qmc(self.h, C.byref(q))
        if q.rng_kind == 0:
            use_ref_lib()

    def init(self):
        self._chk(self.L.orc_init(self.h))

    def run(self):
        self._chk(self.L.orc_run(self.h))
        return self.rows()

    def run_ccmc(self):
        """ccmc{...}: returns (rows, nattempts column)."""
        self._chk(self.L.orc_run_ccmc(self.h))
        rows = self.rows()
        na = np.zeros(len(rows), dtype=np.int64)
        self.L.orc_get_nattempts_rows(self.h, _p(na))
        return rows, na

    def rows(self):
        n = self.L.orc_nrows(self.h)
        out = np.zeros((n, 8))
        self.L.orc_get_rows(self.h, _p(out))
        return out

    def reference(self):
        ref = np.zeros(3)
        occ0 = np.zeros(self.nel, dtype=np.int32)
        f0 = np.zeros(self.W, dtype=np.uint64)
        self.L.orc_reference(self.h, _p(ref), _p(occ0), _p(f0))
        return {"H00": ref[0], "pattempt_single": ref[1], "pattempt_double": ref[2], "occ": occ0, "f0": f0}

    def get_psips(self, rank=0):
        n = self.L.orc_nstates(self.h, rank)
        states = np.zeros((n, self.W), dtype=np.uint64)
        pops = np.zeros(n, dtype=np.int64)
        dat = np.zeros(n)
        self.L.orc_get_psips(self.h, rank, _p(states), _p(pops), _p(dat))
        return states, pops, dat

    def set_psips(self, states, pops, dat, rank=0):
        states = np.ascontiguousarray(states, dtype=np.uint64)
        pops = np.ascontiguousarray(pops, dtype=np.int64)
        dat = np.ascontiguousarray(dat, dtype=np.float64)
        self.L.orc_set_psips(self.h, rank, len(pops), _p(states), _p(pops), _p(dat))

    def set_reference_det(self, f0):
        f0 = np.ascontiguousarray(f0, dtype=np.uint64)
        self.L.orc_set_reference_det(self.h, _p(f0))

    def iterate(self, ncycles, first_cycle, tau, shift, proj_energy_old):
        out = np.zeros(16)
        self._chk(self.L.orc_iterate(self.h, ncycles, first_cycle, tau, shift, proj_energy_old, _p(out)))
        keys = ["proj_energy", "D0_population", "nparticles", "nstates", "nspawn_events", "ndeath", "rspawn", "error",
                "nattempts", "ndraws"]
        return dict(zip(keys, out))

    def stage_spawn(self, cycle, tau, shift, proj_energy_old, rank=0):
        """Spawn/death loop of one cycle only; returns (stats, sdata[n][W+2]) of `rank` before the exchange."""
        out = np.zeros(4)
        self._chk(self.L.orc_stage_spawn(self.h, cycle, tau, shift, proj_energy_old, _p(out)))
        n = self.L.orc_spawn_count(self.h, rank)
        sd = np.zeros((n, self.W + 2), dtype=np.int64)
        self.L.orc_get_spawn(self.h, rank, _p(sd))
        return dict(zip(["proj_energy", "D0_population", "nspawn_events", "ndeath"], out)), sd

    def ccmc_stage_spawn(self, cycle, tau, shift, proj_energy_old):
        """CCMC cluster selection / spawning / death of one cycle (rank 0); returns (stats, sdata[n][W+2])."""
        out = np.zeros(8)
        self._chk(self.L.orc_ccmc_stage_spawn(self.h, cycle, tau, shift, proj_energy_old, _p(out)))
        n = self.L.orc_spawn_count(self.h, 0)
        sd = np.zeros((n, self.W + 2), dtype=np.int64)
        self.L.orc_get_spawn(self.h, 0, _p(sd))
        keys = ["proj_energy", "D0_population", "D0_normalisation", "nattempts", "nattempts_spawn", "nspawn_events",
                "ndeath", "ndeath_nc"]
        return dict(zip(keys, out)), sd

    def ccmc_set_full_nc(self, full_nc=True):
        """ccmc = { full_non_composite = true }"""
        self.L.orc_ccmc_set_full_nc(self.h, int(full_nc))

    def set_pattempt_update(self, on=True):
        """qmc = { pattempt_update = true }: pattempt_single follows the spawn statistics until the shift varies"""
        self.L.orc_set_pattempt_update(self.h, int(on))

    ccmc_set_pattempt_update = set_pattempt_update

    def pattempt_log(self):
        out = np.zeros(4096)
        n = int(self.L.orc_get_pattempt_log(self.h, out.ctypes.data_as(C.c_void_p), len(out)))
        return out[:min(n, len(out))]

    ccmc_pattempt_log = pattempt_log

    def ps_stats(self, rank=0, reset=False):
        """rep_accum of one rank: h_pgen_singles_sum, excit_gen_singles, h_pgen_doubles_sum, excit_gen_doubles"""
        out = np.zeros(4)
        self.L.orc_get_ps_stats(self.h, int(rank), out.ctypes.data_as(C.c_void_p), int(reset))
        return out

    def set_semi_stoch(self, space="high", size=0, start_iteration=1, shift_start_iteration=-1, ci_ex_level=-1,
                       separate_annihilation=True, pop_real_bits=31):
        """semi_stoch = { space = "high" | "ci", size, start_iteration, ... } (before init); pop_real_bits = 11 restates
        real_amplitude_force_32"""
        if shift_start_iteration != -1:          # read_semi_stoch_in (src/lua_hande_calc.f90:1712-1715)
            start_iteration = 2**31 - 1
        self.L.orc_set_semi_stoch.argtypes = [C.c_void_p] + [C.c_int] * 7
        self.L.orc_set_semi_stoch(self.h, {"none": 0, "high": 1, "ci": 2}[space], int(size), int(start_iteration),
                                  int(shift_start_iteration), int(ci_ex_level), int(separate_annihilation), int(pop_real_bits))

    def set_vary_shift(self, on=True):
        """qmc = { vary_shift = true } (after init)"""
        self.L.orc_set_vary_shift.argtypes = [C.c_void_p, C.c_int]
        self.L.orc_set_vary_shift(self.h, int(on))

    def init_semi_stoch(self, dets=None, sizes=None):
        """init_semi_stoch_t on the current lists; with dets (ntot x W, rank by rank) and sizes the space is the caller's"""
        if dets is None:
            self._chk(self.L.orc_init_semi_stoch(self.h))
            return
        dets = np.ascontiguousarray(dets, dtype=np.uint64)
        sizes = np.ascontiguousarray(sizes, dtype=np.int32)
        self.L.orc_init_semi_stoch_dets.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        self._chk(self.L.orc_init_semi_stoch_dets(self.h, _p(dets), _p(sizes)))

    def determ_space(self):
        """(dets [tot x W], sizes per rank)"""
        sizes = np.zeros(self.qmc.nprocs, dtype=np.int32)
        self.L.orc_determ_sizes.argtypes = [C.c_void_p, C.c_void_p]
        tot = self.L.orc_determ_sizes(self.h, _p(sizes))
        dets = np.zeros((tot, self.W), dtype=np.uint64)
        self.L.orc_get_determ_dets.argtypes = [C.c_void_p, C.c_void_p]
        if tot:
            self.L.orc_get_determ_dets(self.h, _p(dets))
        return dets, sizes

    def determ_hamil(self, rank=0):
        """determ%hamil of one rank as (row_ptr, col_ind, mat), 0-based"""
        self.L.orc_get_determ_hamil.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        self.L.orc_get_determ_hamil.restype = C.c_int64
        nnz = self.L.orc_get_determ_hamil(self.h, int(rank), None, None, None)
        dets, sizes = self.determ_space()
        rp = np.zeros(len(dets) + 1, dtype=np.int32)
        ci = np.zeros(max(nnz, 1), dtype=np.int32)
        mat = np.zeros(max(nnz, 1))
        self.L.orc_get_determ_hamil(self.h, int(rank), _p(rp), _p(ci), _p(mat))
        return rp, ci[:nnz], mat[:nnz]

    def determ_vector(self, rank=0):
        dets, sizes = self.determ_space()
        v = np.zeros(int(sizes[rank]))
        self.L.orc_nstates.restype = C.c_int64
        self.L.orc_nstates.argtypes = [C.c_void_p, C.c_int]
        fl = np.zeros(int(self.L.orc_nstates(self.h, int(rank))), dtype=np.uint8)
        self.L.orc_get_determ_vector.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        self.L.orc_get_determ_vector(self.h, int(rank), _p(v) if len(v) else None, _p(fl) if len(fl) else None)
        return v, fl

    def set_quasi_newton(self, on=True, threshold=-1.0, value=-1.0, pop_control=-1.0):
        """qmc = { quasi_newton = true, ... } (before init)"""
        self.L.orc_set_quasi_newton.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double]
        self.L.orc_set_quasi_newton(self.h, int(on), float(threshold), float(value), float(pop_control))

    def quasi_newton(self):
        out, sp = np.zeros(4), np.zeros(self.nbasis + 1)
        self.L.orc_get_quasi_newton.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        self.L.orc_get_quasi_newton(self.h, out.ctypes.data_as(C.c_void_p), sp.ctypes.data_as(C.c_void_p))
        return dict(ref_fock_sum=out[0], threshold=out[1], value=out[2], pop_control=out[3], sp_fock=sp)

    def pattempt_parallel(self):
        self.L.orc_get_pattempt_parallel.restype = C.c_double
        self.L.orc_get_pattempt_parallel.argtypes = [C.c_void_p]
        return float(self.L.orc_get_pattempt_parallel(self.h))

    def set_pattempt_parallel(self, pp):
        self.L.orc_set_pattempt_parallel.argtypes = [C.c_void_p, C.c_double]
        self.L.orc_set_pattempt_parallel(self.h, float(pp))

    def set_pattempt(self, ps, pd):
        self.L.orc_set_pattempt.argtypes = [C.c_void_p, C.c_double, C.c_double]
        self.L.orc_set_pattempt(self.h, float(ps), float(pd))

    def ccmc_hash_shift(self):
        return int(self.L.orc_ccmc_get_hash_shift(self.h))

    def stage_annihilate(self):
        self._chk(self.L.orc_stage_annihilate(self.h))

    # ---- wall-Chebyshev propagator (src/propagators.f90)
    def set_harmonic_forcing(self, value):
        """qmc = { shift_harmonic_forcing = value } (after init)"""
        self.L.orc_set_harmonic_forcing.argtypes = [C.c_void_p, C.c_double]
        self.L.orc_set_harmonic_forcing(self.h, float(value))

    def init_chebyshev(self, order=5, shift=0.0, scale=1.1, skip_gershgorin=False, harmonic_forcing=0.0):
        """after init(); returns (upper spectral bound, zeroes, weights).  The oracle sets tau itself? no: as in the
        reference the caller passes tau = 1 (lua_hande_calc.f90:1410)."""
        out = np.zeros(1 + 2 * order)
        self.L.orc_init_chebyshev.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, C.c_void_p]
        self._chk(self.L.orc_init_chebyshev(self.h, order, shift, scale, int(skip_gershgorin), harmonic_forcing, _p(out)))
        return out[0], out[1:1 + order].copy(), out[1 + order:].copy()

    def set_propagator_weight(self, w):
        self.L.orc_set_propagator_weight.argtypes = [C.c_void_p, C.c_double]
        self.L.orc_set_propagator_weight(self.h, float(w))

    def set_chebyshev_step(self, icheb, shift=None):
        """select the sub-cycle whose weight the next staged cycle uses; shift != None: update_chebyshev(shift) first"""
        self.L.orc_set_chebyshev_step.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int]
        self._chk(self.L.orc_set_chebyshev_step(self.h, icheb, 0.0 if shift is None else shift, int(shift is not None)))

    # ---- load balancing (src/load_balancing.F90, src/qmc_common.F90:505-595, 1332-1390)
    def proc_map(self):
        out = np.zeros(4096, dtype=np.int32)
        n = self.L.orc_get_proc_map(self.h, _p(out))
        return out[:n].copy()

    def set_proc_map(self, pmap):
        pmap = np.ascontiguousarray(pmap, dtype=np.int32)
        self._chk(self.L.orc_set_proc_map(self.h, _p(pmap), len(pmap)))

    def slot_pop(self, rank=0):
        out = np.zeros(4096)
        n = self.L.orc_slot_pop(self.h, rank, _p(out))
        return out[:n].copy()

    def do_load_balancing(self, percent=0.05):
        r = self.L.orc_do_load_balancing(self.h, C.c_double(percent))
        self._chk(min(r, 0))
        return bool(r)

    def redistribute(self, cycle):
        self.L.orc_redistribute.argtypes = [C.c_void_p, C.c_uint32]
        self._chk(self.L.orc_redistribute(self.h, cycle))

    def gen_excit_philox(self, f, cycle, attempt, parent_pop, tau):
        f = np.ascontiguousarray(f, dtype=np.uint64)
        io = np.zeros(8, dtype=np.int32)
        do = np.zeros(2)
        ns = np.zeros(1, dtype=np.int64)
        self._chk(self.L.orc_gen_excit_philox(self.h, _p(f), cycle, attempt, parent_pop, tau, _p(io), _p(do), _p(ns)))
        return io, do, int(ns[0])

    def gen_excit_list(self, f, rn):
        f = np.ascontiguousarray(f, dtype=np.uint64)
        rn = np.ascontiguousarray(rn, dtype=np.float64)
        io = np.zeros(8, dtype=np.int32)
        do = np.zeros(2)
        k = self._chk(self.L.orc_gen_excit_list(self.h, _p(f), _p(rn), len(rn), _p(io), _p(do)))
        return io, do, k

    def gen_excit_spawn_list(self, f, parent_pop, tau, rn):
        """gen_excit + attempt_to_spawn on an injected list of uniform numbers; returns (io, do, nspawn, numbers used)"""
        f = np.ascontiguousarray(f, dtype=np.uint64)
        rn = np.ascontiguousarray(rn, dtype=np.float64)
        io = np.zeros(8, dtype=np.int32)
        do = np.zeros(2)
        ns = np.zeros(1, dtype=np.int64)
        self.L.orc_gen_excit_spawn_list.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_void_p, C.c_int,
                                                    C.c_void_p, C.c_void_p, C.c_void_p]
        k = self._chk(self.L.orc_gen_excit_spawn_list(self.h, _p(f), int(parent_pop), tau, _p(rn), len(rn), _p(io), _p(do), _p(ns)))
        return io, do, int(ns[0]), k

    def heat_bath_tables(self):
        nb = self.L.orc_hb_nb(self.h)
        shp = {0: (nb,), 1: (nb, nb), 2: (nb,) * 3, 3: (nb,) * 3, 4: (nb, nb), 5: (nb,) * 4, 6: (nb,) * 4,
               7: (nb,) * 3}
        names = ["i_weights", "ij_weights", "ija_w", "ija_U", "ija_tot", "ijab_w", "ijab_U", "ijab_tot"]
        out = {}
        for k, nm in enumerate(names):
            n = int(np.prod(shp[k]))
            out[nm] = np.ctypeslib.as_array(self.L.orc_hb_ptr_d(self.h, k), shape=(n,))
        out["ija_K"] = np.ctypeslib.as_array(self.L.orc_hb_ptr_i(self.h, 0), shape=(nb ** 3,))
        out["ijab_K"] = np.ctypeslib.as_array(self.L.orc_hb_ptr_i(self.h, 1), shape=(nb ** 4,))
        out["nb"] = nb
        return out

    def power_pitzer_orderN_tables(self):
        """ppn_* alias tables of excit_gen = power_pitzer_orderN: list of 6 dicts (i_s, ia_s, i_d, ij_d, ia_d, jb_d)"""
        L = self.L
        L.orc_ppn_ptr_d.restype = C.POINTER(C.c_double)
        L.orc_ppn_ptr_d.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.orc_ppn_ptr_i.restype = C.POINTER(C.c_int)
        L.orc_ppn_ptr_i.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_ppn_occ.restype = C.POINTER(C.c_int)
        L.orc_ppn_occ.argtypes = [C.c_void_p]
        out = []
        n = C.c_int64(0)
        for which in range(6):
            t = {}
            for part, nm in enumerate(("w", "U", "tot")):
                ptr = L.orc_ppn_ptr_d(self.h, which, part, C.byref(n))
                t[nm] = np.ctypeslib.as_array(ptr, shape=(n.value,)).copy()
            ptr = L.orc_ppn_ptr_i(self.h, which, C.byref(n))
            t["K"] = np.ctypeslib.as_array(ptr, shape=(n.value,)).astype(np.int32)
            out.append(t)
        nel = len(out[0]["w"])
        occ = np.ctypeslib.as_array(L.orc_ppn_occ(self.h), shape=(nel,)).astype(np.int32)
        return out, occ

    def power_pitzer_tables(self):
        """pp_ia_d / pp_jb_d alias tables and virtual lists of excit_gen = power_pitzer (reference-mapped; on the UEG
        only pp_ia_d exists, one column per orbital)"""
        L = self.L
        L.orc_pp_ptr_d.restype = C.POINTER(C.c_double)
        L.orc_pp_ptr_d.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.orc_pp_ptr_i.restype = C.POINTER(C.c_int)
        L.orc_pp_ptr_i.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_pp_virt.restype = C.POINTER(C.c_int)
        L.orc_pp_virt.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_ppn_occ.restype = C.POINTER(C.c_int)
        L.orc_ppn_occ.argtypes = [C.c_void_p]

        def arr(ptr, n, dtype):
            return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype).copy() if n > 0 else np.zeros(0, dtype=dtype)
        out = []
        n = C.c_int64(0)
        for which in range(2):
            t = {}
            for part, nm in enumerate(("w", "U", "tot")):
                ptr = L.orc_pp_ptr_d(self.h, which, part, C.byref(n))
                t[nm] = arr(ptr, n.value, np.float64)
            ptr = L.orc_pp_ptr_i(self.h, which, C.byref(n))
            t["K"] = arr(ptr, n.value, np.int32)
            out.append(t)
        virt = []
        m = C.c_int(0)
        for spin in range(2):
            ptr = L.orc_pp_virt(self.h, spin, C.byref(m))
            virt.append(arr(ptr, m.value, np.int32))
        nocc = len(out[1]["tot"]) and len(out[0]["tot"])
        occ = arr(L.orc_ppn_occ(self.h), nocc, np.int32)
        return out, virt, occ, int(L.orc_pp_stride(self.h))

    def cpu_baseline(self, nthreads, ncycles, tau, shift, proj_energy_old):
        out = np.zeros(4)
        self._chk(self.L.orc_cpu_baseline(self.h, nthreads, ncycles, tau, shift, proj_energy_old, _p(out)))
        return {"seconds": out[0], "walker_iters": out[1], "attempts": out[2]}


def dsfmt_stream(seed, n):
    use_ref_lib()
    out = np.zeros(n)
    lib().orc_dsfmt_stream(seed, n, _p(out))
    return out


def philox_stream(seed, cycle, purpose, f, attempt, n):
    f = np.ascontiguousarray(f, dtype=np.uint64)
    out = np.zeros(n)
    lib().orc_philox_stream(seed, cycle, purpose, _p(f), len(f), attempt, n, _p(out))
    return out
