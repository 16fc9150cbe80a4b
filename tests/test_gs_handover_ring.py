"""CPU model of the hand-over ring of gs_grid_kernel (csrc/factor.cu).  A finished row travels as
tagged words in ring slot (event mod nslot); the host sizes the ring with
gs_ll_slots(G) = (GS_MAX_REITER + 1) * (G + 2).  A slot may be overwritten only when every CTA has
consumed the event it held.  The kernel has no back-pressure; what protects a slot is the order of
the protocol itself: a CTA publishes an event only after it has consumed every earlier one, and it
owns one row in any G consecutive rows (rows are dealt round-robin).  This test drives that
protocol with adversarial random schedules (any CTA that can make progress may be the one that
does, the others stall arbitrarily long) and random re-iteration counts, and checks that a
publisher never runs a full ring ahead of the slowest consumer."""
import random

import pytest

GS_MAX_REITER = 3


def gs_ll_slots(G):
    return (GS_MAX_REITER + 1) * (G + 2)


def worst_lag(G, rows, seed, slow=None):
    rnd = random.Random(seed)
    # events of row i: 0..3 re-iterations (decision REITER) then one FINAL / REMOVED decision
    events = []                                   # (row, owner)
    for i in range(rows):
        for _ in range(rnd.choice([0, 0, 0, 1, 1, 2, 3]) + 1):
            events.append((i, i % G))
    E = len(events)
    consumed = [0] * G                            # events consumed (the owner consumes its own at once)
    published = 0
    worst = 0
    while min(consumed) < E:
        movers = []
        for c in range(G):
            if consumed[c] < published:
                movers.append(("consume", c))
        if published < E:
            o = events[published][1]
            if consumed[o] == published:          # the owner has seen everything before: it may publish
                movers.append(("publish", o))
        kind, c = rnd.choice(movers) if slow is None or rnd.random() < 0.05 or all(m[1] == slow for m in movers) \
            else rnd.choice([m for m in movers if m[1] != slow])
        if kind == "consume":
            consumed[c] += 1
        else:
            # the slot of event `published` still holds event published - nslot: everybody must be past it
            worst = max(worst, published - min(consumed))
            published += 1
            consumed[c] += 1
    return worst


@pytest.mark.parametrize("G", [2, 8, 25, 64, 148])
def test_publisher_never_laps_the_slowest_consumer(G):
    nslot = gs_ll_slots(G)
    for seed in range(3):
        # uniformly random schedules, and schedules in which one CTA almost never runs
        assert worst_lag(G, 4 * G + 3, seed) < nslot
        assert worst_lag(G, 4 * G + 3, seed, slow=seed % G) < nslot


def _warp_reduce_multi(vals):
    """Python replay of warp_reduce_multi<N> (csrc/factor.cu): vals[lane][n]; returns v[0] per lane."""
    N = len(vals[0])
    lg = N.bit_length() - 1
    v = [list(x) for x in vals]
    for s in range(lg):
        M, o = N >> s, 16 >> s
        new = [list(x) for x in v]
        for lane in range(32):
            upper = (lane & o) != 0
            for j in range(M // 2):
                send_partner = v[lane ^ o][j] if (((lane ^ o) & o) != 0) else v[lane ^ o][j + M // 2]
                keep = v[lane][j + M // 2] if upper else v[lane][j]
                new[lane][j] = keep + send_partner
        v = new
    o = 16 >> lg
    while o:
        v = [[v[lane][0] + v[lane ^ o][0]] + v[lane][1:] for lane in range(32)]
        o >>= 1
    return [x[0] for x in v]


@pytest.mark.parametrize("N", [1, 2, 4, 8, 16])
def test_multi_value_warp_reduction_lands_sum_n_on_lane_group_n(N):
    """After the exchange levels lane l holds the warp sum of the ORIGINAL value l >> (5 - log2 N):
    the layout block_sum8 relies on when it writes the warp partials."""
    rnd = random.Random(N)
    vals = [[rnd.randint(-1000, 1000) for _ in range(N)] for _ in range(32)]
    out = _warp_reduce_multi(vals)
    sh = 5 - (N.bit_length() - 1)
    for lane in range(32):
        assert out[lane] == sum(vals[l][lane >> sh] for l in range(32))
