"""A small executable MODEL of the epoch-flag protocol of csrc/comm.cu + amt_pipe.cu (the fused halo exchange), checked
under thousands of random interleavings and, for two ranks, exhaustively.  It does not run the CUDA code (the GPU
tests do that, bit for bit); it checks the DESIGN: with every rank issuing the same stream-ordered sequence

    [stand-in advance_uv of step n]  ->  [u push kernel]  ->  advance_mu_t of step n = { south-row blocks: v push,
                                                              east-column / north-row blocks: wait, read halo,
                                                              store outputs, signal; interior blocks }

and kernels of one rank running strictly one after the other while the blocks INSIDE a kernel, and the ranks, run
in any order, the flags guarantee that
  * a halo cell is never read before the neighbour's value of THAT step has arrived (read-after-write),
  * a halo cell is never overwritten before its reader of the previous step has read it (write-after-read),
  * no schedule deadlocks.
Flag values are step numbers, exactly as in the library: uv_from_{east,north} = n once the neighbour's u / v of step n
is in my halo; out_from_{west,south} = n once the west / south neighbour's edge blocks have finished step n (its mu,
muts, mudf edges are in my halo and it no longer reads the u / v halo I filled for step n)."""
import itertools
import random

import pytest

W, E, S, N = "W", "E", "S", "N"


class Rank:
    def __init__(self, r, px, py):
        self.r, self.pi, self.pj = r, r % px, r // px
        self.nbr = {W: r - 1 if self.pi > 0 else None, E: r + 1 if self.pi + 1 < px else None,
                    S: r - px if self.pj > 0 else None, N: r + px if self.pj + 1 < py else None}
        # flags in MY memory (written by neighbours)
        self.uv_from = {E: 0, N: 0}
        self.out_from = {W: 0, S: 0}
        # halo contents: the step whose value currently sits in the cell (0 = initial upload)
        self.uv_halo = {E: 0, N: 0}          # u east halo / v north halo
        self.out_halo = {W: 0, S: 0}         # mudf west / south halo
        self.kernels = []                    # stream: list of kernels; a kernel = list of pending block groups
        self.pos = 0                         # index of the running kernel


def build(px, py, nsteps, standin):
    ranks = [Rank(r, px, py) for r in range(px * py)]
    for rk in ranks:
        for n in range(1, nsteps + 1):
            if standin and n > 1:
                rk.kernels.append([("standin", n)])
            if rk.nbr[W] is not None:
                rk.kernels.append([("push_u", n)])
            groups = [("interior", n)]
            if rk.nbr[S] is not None:
                groups.append(("south_row", n))
            if rk.nbr[E] is not None:
                groups.append(("east_col", n))
            if rk.nbr[N] is not None:
                groups.append(("north_row", n))
            rk.kernels.append(groups)
    return ranks


def enabled(ranks, rk, ev):
    """Can block group `ev` of rank `rk` pass its flag wait now?"""
    kind, n = ev
    if kind == "interior":
        return True
    if kind == "standin":          # reads the west / south halos of mudf written by step n-1
        return all(rk.out_from[s] >= n - 1 for s in (W, S) if rk.nbr[s] is not None)
    if kind == "push_u":           # write-after-read: the west neighbour's east-column blocks finished step n-1
        return rk.out_from[W] >= n - 1
    if kind == "south_row":        # likewise for the south neighbour's north-row blocks
        return rk.out_from[S] >= n - 1
    if kind == "east_col":
        return rk.uv_from[E] >= n
    if kind == "north_row":
        return rk.uv_from[N] >= n
    raise AssertionError(kind)


def execute(ranks, rk, ev):
    kind, n = ev
    if kind == "standin":
        for s in (W, S):
            if rk.nbr[s] is not None:
                assert rk.out_halo[s] == n - 1, f"rank {rk.r}: stand-in of step {n} read mudf halo of step {rk.out_halo[s]}"
    elif kind == "push_u":
        w = ranks[rk.nbr[W]]
        assert w.uv_halo[E] == n - 1, f"u halo of rank {w.r} overwritten out of order"
        w.uv_halo[E] = n
        w.uv_from[E] = n
    elif kind == "south_row":
        s = ranks[rk.nbr[S]]
        assert s.uv_halo[N] == n - 1
        s.uv_halo[N] = n
        s.uv_from[N] = n
    elif kind in ("east_col", "north_row"):
        side = E if kind == "east_col" else N
        assert rk.uv_halo[side] == n, f"rank {rk.r}: step {n} read the {side} halo of step {rk.uv_halo[side]}"
        nb = ranks[rk.nbr[side]]
        back = W if side == E else S
        assert nb.out_halo[back] == n - 1
        nb.out_halo[back] = n                 # mu, muts, mudf edge of step n
        nb.out_from[back] = n                 # ... and "I have finished reading your halo of step n"


def frontier(ranks):
    """All (rank, block group) pairs that could run next: groups of each rank's CURRENT kernel only."""
    out = []
    for rk in ranks:
        if rk.pos < len(rk.kernels):
            for ev in rk.kernels[rk.pos]:
                if enabled(ranks, rk, ev):
                    out.append((rk, ev))
    return out


def step(ranks, rk, ev):
    execute(ranks, rk, ev)
    rk.kernels[rk.pos].remove(ev)
    if not rk.kernels[rk.pos]:
        rk.pos += 1


def finished(ranks):
    return all(rk.pos == len(rk.kernels) for rk in ranks)


@pytest.mark.parametrize("px,py", [(1, 2), (2, 1), (1, 4), (2, 2), (2, 4), (4, 2), (1, 8)])
@pytest.mark.parametrize("standin", [False, True])
def test_random_interleavings_are_safe_and_live(px, py, standin):
    rng = random.Random(1000 * px + 10 * py + standin)
    for _ in range(150):
        ranks = build(px, py, nsteps=5, standin=standin)
        # adversarial schedulers: uniformly random, or biased towards one rank running far ahead
        favourite = rng.randrange(px * py) if rng.random() < 0.5 else None
        while not finished(ranks):
            f = frontier(ranks)
            assert f, "deadlock: no block group can make progress"
            fav = [x for x in f if x[0].r == favourite]
            rk, ev = rng.choice(fav) if fav and rng.random() < 0.8 else rng.choice(f)
            step(ranks, rk, ev)


def test_two_ranks_exhaustively():
    """Every interleaving of two ranks over three steps (with the stand-in): a few thousand schedules."""
    def explore(ranks, depth=0):
        if finished(ranks):
            return 1
        f = frontier(ranks)
        assert f, "deadlock"
        total = 0
        for idx in range(len(f)):
            import copy
            clone = copy.deepcopy(ranks)
            g = frontier(clone)
            step(clone, g[idx][0], g[idx][1])
            total += explore(clone, depth + 1)
        return total
    for px, py in ((1, 2), (2, 1)):
        assert explore(build(px, py, nsteps=3, standin=True)) > 100


def test_the_model_catches_a_missing_guard():
    """Sanity of the checker itself: drop the write-after-read wait of the v push and some schedule must fail."""
    global enabled
    orig = enabled

    def broken(ranks, rk, ev):
        return True if ev[0] == "south_row" else orig(ranks, rk, ev)
    enabled = broken
    try:
        rng = random.Random(7)
        failed = False
        for _ in range(300):
            ranks = build(1, 3, nsteps=4, standin=False)
            try:
                while not finished(ranks):
                    f = frontier(ranks)
                    if not f:
                        failed = True
                        break
                    rk, ev = rng.choice(f)
                    step(ranks, rk, ev)
            except AssertionError:
                failed = True
            if failed:
                break
        assert failed, "the model did not notice a missing write-after-read guard"
    finally:
        enabled = orig
