"""Host side of load balancing: the policy of `do_load_balancing` (src/load_balancing.F90:209-323) on the slot
populations the engine reports (hb200_slot_populations = initialise_slot_pop).  SURVEY.md 8: the policy stays on the
host; the engine consumes its output (hb200_set_proc_map) and moves the determinants (hb200_redistribute_particles).
Same routines, same order of operations and the same tie rules as the reference, so that every rank (and the oracle)
derives the same proc_map from the same all-reduced slot list."""
import numpy as np


def check_imbalance(nparticles_proc, average_pop, percent_imbal):
    """src/load_balancing.F90:353-381"""
    return bool((np.asarray(nparticles_proc) > average_pop + average_pop * percent_imbal).any())


def insertion_rank(arr, tol=0.0):
    """lib/local/ranking.f90:81-117 (0-based ranks): rank[k] = index of the k-th smallest entry, entries closer than
    `tol` keep their order."""
    arr = np.asarray(arr, dtype=np.float64)
    rank = list(range(len(arr)))
    for i in range(1, len(arr)):
        j = i - 1
        tmp = rank[i]
        while j >= 0:
            if arr[rank[j]] - arr[tmp] < tol:
                break
            rank[j + 1] = rank[j]
            j -= 1
        rank[j + 1] = tmp
    return rank


def find_processors(procs_pop, up_thresh, low_thresh, proc_map):
    """src/load_balancing.F90:520-601: receivers (the nrecv least populated ranks, ascending), donors, number of donor
    slots."""
    nrecv = sum(1 for x in procs_pop if x < low_thresh)
    donors = [i for i, x in enumerate(procs_pop) if not (x < low_thresh) and x > up_thresh]
    rank = insertion_rank(procs_pop, 1.0e-8)
    receivers = rank[:nrecv]
    donor_slots = sum(1 for m in proc_map for d in donors if m == d)
    return receivers, donors, donor_slots


def reduce_slots(donors, slot_list, proc_map):
    """src/load_balancing.F90:478-518"""
    idx, pop = [], []
    for d in donors:
        for j in range(len(slot_list)):
            if proc_map[j] == d:
                pop.append(slot_list[j]); idx.append(j)
    return idx, pop


def redistribute_slots(d_index, d_pop, d_rank, receivers, up_thresh, low_thresh, proc_map, procs_pop):
    """src/load_balancing.F90:419-476: smallest donor slots first, first receiver that stays below the upper threshold
    while the donor stays above the lower one."""
    for pos in d_rank:
        for r in receivers:
            new_pop = d_pop[pos] + procs_pop[r]
            donor_pop = procs_pop[proc_map[d_index[pos]]] - d_pop[pos]
            if donor_pop >= low_thresh and new_pop <= up_thresh:
                procs_pop[proc_map[d_index[pos]]] = donor_pop
                procs_pop[r] = new_pop
                proc_map[d_index[pos]] = r
                break


def do_load_balancing(slot_list, proc_map, nprocs, percent=0.05):
    """do_load_balancing given the all-reduced slot populations.  Returns (needed, new proc_map, nparticles_proc)."""
    slot_list = np.asarray(slot_list, dtype=np.float64)
    proc_map = [int(x) for x in proc_map]
    procs_pop = [float(sum(slot_list[j] for j in range(len(proc_map)) if proc_map[j] == r)) for r in range(nprocs)]
    pop_av = sum(procs_pop) / nprocs
    if not check_imbalance(procs_pop, pop_av, percent):
        return False, proc_map, procs_pop
    up_thresh = pop_av + int(pop_av * percent)
    low_thresh = pop_av - int(pop_av * percent)
    receivers, donors, _ = find_processors(procs_pop, up_thresh, low_thresh, proc_map)
    d_index, d_pop = reduce_slots(donors, slot_list, proc_map)
    d_rank = insertion_rank(d_pop, 1.0e-8)
    redistribute_slots(d_index, d_pop, d_rank, receivers, up_thresh, low_thresh, proc_map, procs_pop)
    return True, proc_map, procs_pop
