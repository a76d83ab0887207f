"""Host side of the wall-Chebyshev propagator (`qmc = { chebyshev = { chebyshev_order = m, ... } }`): `init_chebyshev`,
`highest_det` and `update_chebyshev` of src/propagators.f90:11-208.  The engine applies the weight of the current
sub-cycle (hb200_set_propagator_weight); the spectral range, the zeroes S_i and the weights 1/(S_i - E_0) live here, as
they do on the Fortran host.  As in the reference the wall-Chebyshev projector runs with tau = 1
(src/lua_hande_calc.f90:1403-1411)."""
import math

PI = 3.1415926535897931


def highest_det(sys):
    """src/propagators.f90:167-186"""
    return [sys.nbasis - (2 * ia - 1) for ia in range(1, sys.nalpha + 1)] + \
           [sys.nbasis - 2 * (ib - 1) for ib in range(1, sys.nbeta + 1)]


def _abs_single(sys, occ, i, a):
    """|<D|H|D_i^a>| (slater_condon1_mol, src/hamiltonian_molecular.f90:141-197); the sign plays no role here"""
    if sys.sym[i] != sys.sym[a] or sys.ms[i] != sys.ms[a]:
        return 0.0
    h = sys.get_one_body(i, a)
    for j in occ:
        if j != i:
            h = h + sys.get_two_body(i, j, a, j)
            h = h - sys.get_two_body(i, j, j, a)
    return abs(h)


def _abs_double(sys, i, j, a, b):
    """|<D|H|D_ij^ab>| (slater_condon2_mol, src/hamiltonian_molecular.f90:261-298)"""
    return abs(sys.get_two_body(i, j, a, b) - sys.get_two_body(i, j, b, a))


class Chebyshev:
    """cheb_t (src/qmc_data.f90:886-907)"""

    def __init__(self, sys, H00, order=5, shift=0.0, scale=1.1, skip_gershgorin=False):
        self.order = int(order)
        occ_max = highest_det(sys)
        hmm = sys.slater_condon0(sorted(occ_max))
        if not skip_gershgorin:
            # Gershgorin circle of the highest determinant over the determinants of the calculation's symmetry within
            # two excitations of it (enumerate_determinants with ref_sym = sys%symmetry)
            e_max = 0.0
            occ = sorted(occ_max)
            occset = set(occ)
            virt = [o for o in range(1, sys.nbasis + 1) if o not in occset]
            if sys.symmetry_orb_list(occ) == sys.symmetry:
                for ii, i in enumerate(occ):
                    for a in virt:
                        e_max = e_max + _abs_single(sys, occ, i, a)
                        for j in occ[ii + 1:]:
                            for b in virt:
                                if b > a:
                                    e_max = e_max + _abs_double(sys, i, j, a, b)
                e_max = e_max + abs(hmm)
            e_max = e_max - abs(hmm) + hmm - H00
        else:
            e_max = hmm - H00
        e_max = (e_max + shift) * scale
        self.spectral_range = [0.0, e_max]
        self.zeroes = [0.0] * self.order
        self.weights = [1.0] * self.order
        self.update(0.0)

    def update(self, shift):
        """update_chebyshev (src/propagators.f90:188-208)"""
        self.spectral_range[0] = shift
        for i in range(1, self.order + 1):
            self.zeroes[i - 1] = shift + (self.spectral_range[1] - self.spectral_range[0]) / 2 * \
                (1 - math.cos(PI * i / (self.order + 0.5)))
            self.weights[i - 1] = 1 / (self.zeroes[i - 1] - shift)
