"""Host-side mirror of HANDE's `sys = ueg { electrons, ms, dim = 3, cutoff, rs }` (reference src/lua_hande_system.F90,
src/system.f90:486-509, src/basis.f90:258-501, src/ueg.f90:43-140): the plane-wave basis ordered by kinetic energy,
the wavevector -> basis-function lookup and the `ternary_conserve` table of allowed `a` orbitals per k_i + k_j.

These tables are built once on the host (as the Fortran host does in init_system / init_excit_gen) and handed to the
engine through hb200_set_system_ueg; everything per walker runs on the GPU.
"""
from __future__ import annotations

import math

import numpy as np

from .read_in import DEPSILON, insertion_rank

PI = 3.1415926535897931  # lib/local/const.F90


class UegSystem:
    """sys_t for the 3D uniform electron gas (no twist).  Orbital indices are 1-based; odd = alpha."""
    kind = "ueg"

    def __init__(self, electrons, ms=0, rs=1.0, cutoff=3.0, dim=3):
        if dim != 3:
            raise ValueError("only the 3D UEG is supported")
        self.nel, self.Ms = int(electrons), int(ms)
        self.nalpha, self.nbeta = (self.nel + self.Ms) // 2, (self.nel - self.Ms) // 2
        self.rs, self.ecutoff = float(rs), float(cutoff)
        # src/system.f90:505-509
        self.L = self.rs * ((4 * PI * self.nel) / 3) ** (1.0 / 3.0)
        rl = 1.0 / self.L
        nmax = int(math.ceil(math.sqrt(2 * self.ecutoff)))
        ks, es = [], [0.0]
        # init_model_basis_fns (src/basis.f90:383-418): k outermost, i innermost; keep |k|^2/2 <= ecutoff
        for k in range(-nmax, nmax + 1):
            for j in range(-nmax, nmax + 1):
                for i in range(-nmax, nmax + 1):
                    if (i * i + j * j + k * k) / 2 > self.ecutoff:
                        continue
                    kc = [(i + 0.0) * rl, (j + 0.0) * rl, (k + 0.0) * rl]       # calc_kinetic (src/kpoints.f90:9-53)
                    dot = kc[0] * kc[0]
                    dot = dot + kc[1] * kc[1]
                    dot = dot + kc[2] * kc[2]
                    ks.append((i, j, k))
                    es.append(2 * PI * PI * dot)
        nsp = len(ks)
        rank = insertion_rank(es, DEPSILON)                                         # stable, tolerance depsilon
        self.nbasis = 2 * nsp
        self.W = (self.nbasis + 63) // 64
        self.nvirt = self.nbasis - self.nel
        self.nvirt_alpha, self.nvirt_beta = nsp - self.nalpha, nsp - self.nbeta
        self.kvec = np.zeros((self.nbasis + 1, 3), dtype=np.int32)
        self.sp_eigv = np.zeros(self.nbasis + 1)
        self.ms = np.zeros(self.nbasis + 1, dtype=np.int32)
        for i in range(1, nsp + 1):
            src = rank[i]
            for s in range(2):
                o = 2 * i - 1 + s
                self.kvec[o] = ks[src - 1]
                self.sp_eigv[o] = es[src]
                self.ms[o] = 1 if s == 0 else -1
        # init_ueg_indexing (src/ueg.f90:43-85)
        self.kmax = int(math.ceil(math.sqrt(2 * self.ecutoff)))
        nk = 2 * self.kmax + 1
        self.offset_inds = np.array([1, nk, nk * nk], dtype=np.int32)
        self.offset = int(self.offset_inds.sum()) * self.kmax + 1
        self.lookup = np.full(nk ** 3 + 1, -1, dtype=np.int32)
        alpha = np.arange(1, self.nbasis + 1, 2)
        self.lookup[self.kvec[alpha] @ self.offset_inds + self.offset] = alpha
        # init_ternary_conserve (src/ueg.f90:87-140)
        K = 2 * self.kmax
        D = 2 * K + 1
        self.tern_kmax = K
        tern = np.zeros((D, D, D, self.W + 1), dtype=np.uint64)   # [k3, k2, k1, 0:W] == Fortran (0:W, k1, k2, k3)
        g = np.arange(-K, K + 1)
        k3, k2, k1 = np.meshgrid(g, g, g, indexing="ij")
        for a in alpha:
            ka = self.kvec[a]
            d2 = (k1 - ka[0]) ** 2 + (k2 - ka[1]) ** 2 + (k3 - ka[2]) ** 2
            ok = (d2 / 2 - self.ecutoff) < 1.0e-8
            tern[..., 0] += ok.astype(np.uint64)
            tern[..., 1 + (a - 1) // 64] |= np.where(ok, np.uint64(1) << np.uint64((a - 1) % 64), np.uint64(0))
        self.ternary_conserve = np.ascontiguousarray(tern.reshape(-1))

    # --- determinants (same bit layout as read_in systems, src/basis_types.f90:134-185)
    def encode(self, occ):
        f = np.zeros(self.W, dtype=np.uint64)
        for o in occ:
            f[(o - 1) // 64] |= np.uint64(1) << np.uint64((o - 1) % 64)
        return f

    def decode(self, f):
        out = []
        for iw in range(self.W):
            x = int(f[iw])
            while x:
                b = (x & -x).bit_length() - 1
                out.append(iw * 64 + b + 1)
                x &= x - 1
        return out

    def coulomb_int(self, i, a):
        """coulomb_int_ueg_3d (src/ueg.f90:250-280)"""
        q = self.kvec[i] - self.kvec[a]
        return 1.0 / (PI * self.L * int(q @ q))

    def slater_condon0(self, occ):
        """slater_condon0_ueg (src/hamiltonian_ueg.f90:71-99,127-156) - host copy used for H00."""
        spe = 0.0
        for o in occ:
            spe = spe + float(self.sp_eigv[o])
        ex = 0.0
        for k, i in enumerate(occ):
            for j in occ[k + 1:]:
                if i % 2 == j % 2:
                    ex = ex - self.coulomb_int(i, j)
        return spe + ex

    def aufbau_reference(self):
        """Lowest-energy determinant of the requested spin polarisation (reference = {} default)."""
        return sorted([2 * i - 1 for i in range(1, self.nalpha + 1)] + [2 * i for i in range(1, self.nbeta + 1)])
