"""TEST HARNESS: ctypes access to tests/hdcheck/libhdcheck.so (host build of hande_b200/csrc/hb_core.cuh)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libhdcheck.so")
CORE = os.path.join(HERE, "..", "..", "hande_b200", "csrc", "hb_core.cuh")


def build():
    src = os.path.join(HERE, "hdcheck.cpp")
    if (not os.path.exists(LIB) or os.path.getmtime(src) > os.path.getmtime(LIB)
            or os.path.getmtime(CORE) > os.path.getmtime(LIB)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-Wno-unknown-pragmas",
                               "-shared", "-o", LIB, src])


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class HdCheck:
    def __init__(self, sys, excit_gen, ps, pd, tau, shift, pe_old, real_factor, spawn_cutoff, seed, f0, H00,
                 hb=None):
        build()
        L = self.L = C.CDLL(LIB)
        self.sys = sys
        self.keep = []
        ia = lambda a: np.ascontiguousarray(a, dtype=np.int32)  # noqa: E731
        if getattr(sys, "kind", "read_in") == "ueg":
            arrs = [ia(sys.kvec), np.ascontiguousarray(sys.sp_eigv, dtype=np.float64), ia(sys.offset_inds),
                    ia(sys.lookup), np.ascontiguousarray(sys.ternary_conserve, dtype=np.uint64)]
            self.keep.append(arrs)
            L.hd_set_sys_ueg.argtypes = [C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                         C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
            L.hd_set_sys_ueg(sys.nbasis, sys.nel, sys.L, _p(arrs[0]), _p(arrs[1]), sys.kmax, sys.offset, _p(arrs[2]),
                             _p(arrs[3]), sys.tern_kmax, _p(arrs[4]))
            self._finish(excit_gen, ps, pd, tau, shift, pe_old, real_factor, spawn_cutoff, seed, f0, H00)
            return
        sym, ms, sp = ia(sys.sym), ia(sys.ms), ia(sys.spatial)
        nbss, ssbf = ia(sys.nbasis_sym_spin), ia(sys.sym_spin_basis_fns)
        h1 = np.ascontiguousarray(sys.h1, dtype=np.float64)
        v2 = [np.ascontiguousarray(v, dtype=np.float64) for v in sys.v2]
        ptrs = (C.c_void_p * 4)(*[v.ctypes.data for v in v2] + [None] * (4 - len(v2)))
        self.keep += [sym, ms, sp, nbss, ssbf, h1, v2, ptrs]
        L.hd_set_sys.argtypes = [C.c_int] * 14 + [C.c_double] + [C.c_void_p] * 7
        L.hd_set_sys(sys.nbasis, sys.nel, sys.nsym_tot, sys.sym0, sys.sym_max, sys.pg_mask, sys.Lz_mask,
                     sys.Lz_offset, sys.gamma_sym, int(sys.uhf), sys.nvirt, sys.nvirt_alpha, sys.nvirt_beta,
                     sys.max_nbss, sys.Ecore, _p(sym), _p(ms), _p(sp), _p(nbss), _p(ssbf), _p(h1), ptrs)
        if hb is not None:
            L.hd_set_heat_bath.argtypes = [C.c_void_p] * 10
            arrs = [np.ascontiguousarray(hb[k]) for k in ("i_weights", "ij_weights", "ija_w", "ija_U", "ija_K",
                                                          "ija_tot", "ijab_w", "ijab_U", "ijab_K", "ijab_tot")]
            self.keep.append(arrs)
            L.hd_set_heat_bath(*[_p(a) for a in arrs])
        self._finish(excit_gen, ps, pd, tau, shift, pe_old, real_factor, spawn_cutoff, seed, f0, H00)

    def _finish(self, excit_gen, ps, pd, tau, shift, pe_old, real_factor, spawn_cutoff, seed, f0, H00):
        L = self.L
        f0 = np.ascontiguousarray(f0, dtype=np.uint64)
        L.hd_set_params.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int64,
                                    C.c_int64, C.c_uint32, C.c_void_p, C.c_double]
        L.hd_set_params(excit_gen, ps, pd, tau, shift, pe_old, real_factor, spawn_cutoff, seed, _p(f0), H00)
        L.hd_gen_excit_philox.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int64, C.c_void_p, C.c_void_p,
                                          C.c_void_p]
        L.hd_gen_excit_list.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.hd_sc0.restype = C.c_double
        L.hd_sc0.argtypes = [C.c_void_p]
        L.hd_sc1.restype = C.c_double
        L.hd_sc1.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.hd_sc2.restype = C.c_double
        L.hd_sc2.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.hd_murmur.restype = C.c_int32
        L.hd_murmur.argtypes = [C.c_void_p, C.c_uint32]
        L.hd_owner.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.hd_philox_stream.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_int, C.c_uint32, C.c_int,
                                       C.c_void_p]
        L.hd_proj_hmatel.restype = C.c_double
        L.hd_proj_hmatel.argtypes = [C.c_void_p, C.c_void_p]

    def set_power_pitzer_orderN(self, tables, occ):
        self.L.hd_set_ppn.argtypes = [C.c_int] + [C.c_void_p] * 4
        self.L.hd_set_ppn_occ.argtypes = [C.c_void_p]
        self.keep.append((tables, occ))
        for which, t in enumerate(tables):
            self.L.hd_set_ppn(which, _p(t["w"]), _p(t["U"]), _p(t["K"]), _p(t["tot"]))
        self.L.hd_set_ppn_occ(_p(occ))

    def set_ueg_power_pitzer(self, table):
        self.L.hd_set_pp.argtypes = [C.c_int] + [C.c_void_p] * 4
        self.keep.append(table)
        self.L.hd_set_pp(0, _p(table["w"]), _p(table["U"]), _p(table["K"]), _p(table["tot"]))

    def set_power_pitzer(self, tables, virt, occ, stride):
        self.L.hd_set_pp.argtypes = [C.c_int] + [C.c_void_p] * 4
        self.L.hd_set_pp_virt.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]
        self.L.hd_set_ppn_occ.argtypes = [C.c_void_p]
        self.keep.append((tables, virt, occ))
        for which, t in enumerate(tables):
            self.L.hd_set_pp(which, _p(t["w"]), _p(t["U"]), _p(t["K"]), _p(t["tot"]))
        self.L.hd_set_pp_virt(_p(virt[0]), len(virt[0]), _p(virt[1]), len(virt[1]), stride)
        self.L.hd_set_ppn_occ(_p(occ))

    def set_pattempt_parallel(self, pp):
        self.L.hd_set_pattempt_parallel.argtypes = [C.c_double]
        self.L.hd_set_pattempt_parallel(float(pp))

    def gen_excit_philox(self, f, cycle, attempt, parent_pop):
        f = np.ascontiguousarray(f, dtype=np.uint64)
        io = np.zeros(8, dtype=np.int32)
        do = np.zeros(2)
        ns = np.zeros(1, dtype=np.int64)
        self.L.hd_gen_excit_philox(_p(f), cycle, attempt, parent_pop, _p(io), _p(do), _p(ns))
        return io, do, int(ns[0])

    def gen_excit_list(self, f, rn):
        f = np.ascontiguousarray(f, dtype=np.uint64)
        rn = np.ascontiguousarray(rn, dtype=np.float64)
        io = np.zeros(8, dtype=np.int32)
        do = np.zeros(2)
        k = self.L.hd_gen_excit_list(_p(f), _p(rn), len(rn), _p(io), _p(do))
        return io, do, k

    def sc0(self, f):
        f = np.ascontiguousarray(f, dtype=np.uint64)
        return self.L.hd_sc0(_p(f))

    def sc1(self, f, i, a):
        f = np.ascontiguousarray(f, dtype=np.uint64)
        return self.L.hd_sc1(_p(f), i, a)

    def sc2(self, f, i, j, a, b):
        f = np.ascontiguousarray(f, dtype=np.uint64)
        return self.L.hd_sc2(_p(f), i, j, a, b)

    def murmur(self, f, seed=7):
        f = np.ascontiguousarray(f, dtype=np.uint64)
        return self.L.hd_murmur(_p(f), seed)

    def philox_stream(self, seed, cycle, purpose, f, attempt, n):
        f = np.ascontiguousarray(f, dtype=np.uint64)
        out = np.zeros(n)
        self.L.hd_philox_stream(seed, cycle, purpose, _p(f), len(f), attempt, n, _p(out))
        return out

    def hmatel_pair(self, f1, f2):
        f1 = np.ascontiguousarray(f1, dtype=np.uint64)
        f2 = np.ascontiguousarray(f2, dtype=np.uint64)
        self.L.hd_hmatel_pair.restype = C.c_double
        self.L.hd_hmatel_pair.argtypes = [C.c_void_p, C.c_void_p]
        return self.L.hd_hmatel_pair(_p(f1), _p(f2))

    def check_if_determ(self, sorted_dets, f):
        sd = np.ascontiguousarray(sorted_dets, dtype=np.uint64)
        f = np.ascontiguousarray(f, dtype=np.uint64)
        self.L.hd_check_if_determ.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        return bool(self.L.hd_check_if_determ(_p(sd), len(sd), _p(f)))

    def proj_hmatel(self, f):
        f = np.ascontiguousarray(f, dtype=np.uint64)
        r = np.zeros(1, dtype=np.int32)
        h = self.L.hd_proj_hmatel(_p(f), _p(r))
        return h, int(r[0])
