"""TEST INFRASTRUCTURE: an Engine stand-in that runs ONE rank of the oracle per process and exchanges the spawn
blocks through torch.distributed (gloo).  It lets the CPU suite exercise the product's host-side multi-rank driver
(hande_b200.fciqmc.do_fciqmc: report-loop reduction, shift update, initial distribution by hash owner) with
world_size 2 and compare it with the reference's np2 golden table.  Never used by the product."""
import ctypes as C

import numpy as np

from oracle import pyoracle
from oracle.pyoracle import Oracle

_PATH = {}


def make_engine_cls(fcidump_path, sys_kw, rng_kind=0, ueg=None, ref_det=None, quasi_newton=None):
    class OracleRankEngine:
        def __init__(self, sys, *, excit_gen, pattempt_single, pattempt_double, real_amplitudes, spawn_cutoff,
                     initiator_approx, initiator_pop, trunc_level, walker_length, spawned_walker_length, seed, nprocs,
                     iproc, nslots, device, pattempt_parallel=-1.0):
            if nprocs > 1:
                import torch.distributed as dist
                self.dist = dist
            self.rank, self.world = iproc, nprocs
            o = self.o = Oracle()
            if ueg is not None:
                o.init_ueg(*ueg)
                if ref_det is not None:
                    o.set_ref_det(ref_det)
            else:
                o.read_fcidump(fcidump_path, **sys_kw)
            o.set_qmc(seed=seed, excit_gen=excit_gen, rng_kind=rng_kind, real_amplitudes=int(real_amplitudes),
                      spawn_cutoff=spawn_cutoff, initiator_approx=int(initiator_approx), initiator_pop=initiator_pop,
                      ex_level=trunc_level, walker_length=walker_length, spawned_walker_length=spawned_walker_length,
                      nprocs=nprocs, nslots=nslots)
            if pattempt_parallel >= 0:
                o.set_pattempt_parallel(pattempt_parallel)
            if quasi_newton is not None:
                o.set_quasi_newton(True, **quasi_newton)
            o.init()
            L = o.L
            L.orc_rank_spawn.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_double, C.c_double, C.c_double, C.c_void_p]
            L.orc_rank_send_count.restype = C.c_int64
            L.orc_rank_send_count.argtypes = [C.c_void_p, C.c_int, C.c_int]
            L.orc_rank_get_send.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
            L.orc_rank_annihilate.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p]
            self.E = o.W + 2
            self.real_factor = (1 << 31) if real_amplitudes else 1
            self._nparticles = 0.0

        def set_reference(self, f0, H00):
            ref = self.o.reference()
            assert (ref["f0"] == f0).all() and ref["H00"] == H00      # host-side reference == oracle's

        @staticmethod
        def get_unique_id():
            return np.arange(128, dtype=np.uint8)

        def comm_init(self, uid):
            assert (np.asarray(uid) == np.arange(128)).all()           # broadcast reached every rank

        def comm_setup(self, comm, p2p=None):
            uid = self.get_unique_id() if comm.rank == 0 else np.zeros(128, dtype=np.uint8)
            self.comm_init(comm.broadcast_bytes(uid, src=0))
            h = comm.allgather_bytes(np.full(64, comm.rank, dtype=np.uint8))     # the IPC-handle all-gather
            assert h.shape == (comm.size, 64) and all((h[r] == r).all() for r in range(comm.size))

        def set_propagator_weight(self, weight):
            self.o.set_propagator_weight(weight)

        def upload_psips(self, states, pops, dat):
            self.o.set_psips(np.asarray(states).reshape(-1, self.o.W), pops, dat, rank=self.rank)
            self._nparticles = float(np.abs(np.asarray(pops, dtype=np.int64)).sum()) / self.real_factor

        def download_psips(self):
            return self.o.get_psips(self.rank)

        def set_determ_space(self, dets, sizes):
            """init_semi_stoch_t with the driver's space; with one rank per process the amplitudes of all ranks are
            all-gathered between the spawning loop and the annihilation (iterate)"""
            self.o.init_semi_stoch(dets, sizes)
            self.determ_sizes = np.asarray(sizes, dtype=np.int64)
            L = self.o.L
            L.orc_rank_get_dvector.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
            L.orc_rank_determ_project.argtypes = [C.c_void_p, C.c_int, C.c_void_p]

        @property
        def nstates(self):
            return int(self.o.L.orc_nstates(self.o.h, self.rank))

        def iterate(self, ncycles, tau, shift, proj_energy_old, first_cycle):
            o, L = self.o, self.o.L
            out = dict(proj_energy=0.0, D0_population=0.0, rspawn=0.0, nspawn_events=0, ndeath=0, nattempts=0,
                       spawn_error=0, psip_error=0, nattempts_spawn=0, walker_iterations=0.0)
            for c in range(ncycles):
                st = np.zeros(8)
                assert L.orc_rank_spawn(o.h, self.rank, first_cycle + c, tau, shift, proj_energy_old,
                                        st.ctypes.data_as(C.c_void_p)) == 0
                blocks = []
                for d in range(self.world):
                    n = L.orc_rank_send_count(o.h, self.rank, d)
                    buf = np.zeros((n, self.E), dtype=np.int64)
                    L.orc_rank_get_send(o.h, self.rank, d, buf.ctypes.data_as(C.c_void_p))
                    blocks.append(buf)
                if self.world > 1 and getattr(self, "determ_sizes", None) is not None and self.determ_sizes.sum() > 0:
                    # determ_proj_separate_annihil: mpi_allgatherv of determ%vector, then this rank's projection
                    mine = np.zeros(int(self.determ_sizes[self.rank]))
                    if len(mine):
                        L.orc_rank_get_dvector(o.h, self.rank, mine.ctypes.data_as(C.c_void_p))
                    parts = [None] * self.world
                    self.dist.all_gather_object(parts, mine)
                    full = np.ascontiguousarray(np.concatenate(parts))
                    assert L.orc_rank_determ_project(o.h, self.rank, full.ctypes.data_as(C.c_void_p)) == 0
                gathered = [None] * self.world
                if self.world > 1:
                    self.dist.all_gather_object(gathered, blocks)        # comm_spawn_t: personalised all-to-all
                else:
                    gathered = [blocks]
                recv = np.concatenate([gathered[src][self.rank] for src in range(self.world)]).reshape(-1, self.E)
                recv = np.ascontiguousarray(recv)
                res = np.zeros(4)
                assert L.orc_rank_annihilate(o.h, self.rank, recv.ctypes.data_as(C.c_void_p), len(recv),
                                             res.ctypes.data_as(C.c_void_p)) == 0
                out["proj_energy"] += st[0]
                out["D0_population"] += st[1]
                out["nspawn_events"], out["ndeath"], out["nattempts"] = int(st[2]), int(st[3]), int(st[4])
                if st[4] > 0:
                    out["rspawn"] += (st[2] + st[3] / self.real_factor) / st[4]
                out["nparticles"], out["nstates"] = res[0], int(res[1])
                out["spawn_error"] = int(res[2])
            return out

        def ccmc_iterate(self, ncycles, tau, shift, proj_energy_old, first_cycle, ex_level):
            """CCMC cycles on the oracle (single rank): same outputs as Engine.ccmc_iterate."""
            o = self.o
            out = dict(proj_energy=0.0, D0_population=0.0, rspawn=0.0, nspawn_events=0, ndeath=0, nattempts=0,
                       spawn_error=0, psip_error=0, nattempts_spawn=0, walker_iterations=0.0)
            for c in range(ncycles):
                st, _ = o.ccmc_stage_spawn(first_cycle + c, tau, shift, proj_energy_old)
                o.stage_annihilate()
                out["proj_energy"] += st["proj_energy"]
                out["D0_population"] += st["D0_population"]
                out["nspawn_events"], out["ndeath"] = int(st["nspawn_events"]), int(st["ndeath"])
                out["nattempts"] = int(st["nattempts"])
                out["nattempts_spawn"] += int(st["nattempts_spawn"])
                if st["nattempts_spawn"] > 0:
                    out["rspawn"] += st["nspawn_events"] / st["nattempts_spawn"]
            out["nparticles"] = float(o.L.orc_nparticles(o.h, self.rank))
            out["nstates"] = self.nstates
            return out

        def set_quasi_newton(self, sp_fock, ref_fock_sum, threshold, value, pop_control):
            # the stand-in was created with the same options: the host's propagator must equal the oracle's
            q = self.o.quasi_newton()
            assert (np.asarray(sp_fock) == q["sp_fock"]).all() and ref_fock_sum == q["ref_fock_sum"]
            assert (threshold, value, pop_control) == (q["threshold"], q["value"], q["pop_control"])

        def set_pattempt(self, pattempt_single, pattempt_double, accumulate=False):
            self.o.set_pattempt(pattempt_single, pattempt_double)
            self.o.set_pattempt_update(accumulate)

        def get_ps_stats(self, reset=True):
            return self.o.ps_stats(self.rank, reset=reset)

        def last_timing(self):
            return {}

        def close(self):
            pass

    return OracleRankEngine
