"""Host-side check of the retrieval work split (no GPU): the list slot every (cluster, query group) segment writes to --
flat_l2_topk_kernel: slot = cluster - first cluster of the group -- stays below the slot count make_layout sizes the
candidate arrays for, for any number of resident clusters up to 148 / cluster size.  python tools/retr_slots_sim.py"""
import random


def host_slots(n_mgrp, n_tiles, sms, cs):  # make_layout (retrieval.cu)
    W = n_mgrp * n_tiles
    nc = min(max(sms // cs, 1), W)
    per = W // nc
    return (n_tiles + per - 1) // per + 1


def kernel_slots(n_mgrp, n_tiles, ncl):  # the epilogue's segment walk
    W = n_mgrp * n_tiles
    mx = 0
    for c in range(ncl):
        w, we = W * c // ncl, W * (c + 1) // ncl
        while w < we:
            grp = w // n_tiles
            tb = w - grp * n_tiles
            te = min(n_tiles, tb + (we - w))
            w += te - tb
            g0 = grp * n_tiles
            c0 = g0 * ncl // W
            while W * (c0 + 1) // ncl <= g0:
                c0 += 1
            while W * c0 // ncl > g0:
                c0 -= 1
            assert c - c0 >= 0
            mx = max(mx, c - c0)
    return mx + 1


if __name__ == "__main__":
    random.seed(1)
    for it in range(20000):
        cs = random.choice([2, 4, 8])
        n_mgrp, n_tiles = random.randint(1, 40), random.randint(1, random.choice([5, 50, 500, 4000]))
        ncl = random.randint(1, min(148 // cs, n_mgrp * n_tiles))
        h, k = host_slots(n_mgrp, n_tiles, 148, cs), kernel_slots(n_mgrp, n_tiles, ncl)
        assert k <= h, (cs, n_mgrp, n_tiles, ncl, h, k)
    print("20000 random splits: every segment's slot is inside the host's slot count")
