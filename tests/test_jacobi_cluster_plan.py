"""CPU replay of the schedule of csrc/jacobi_cluster.cu (cluster-resident one-sided Jacobi SVD):
the circle-method tournament over 2C blocks as the kernel's pushes implement it (top_i -> top_{i+1},
top_{C-1} -> bot_{C-1}, bot_i -> bot_{i-1}, bot_0 -> top_1, top_0 fixed), the closed form the
kernel uses to name the block a CTA ends up with (jc_block_at), the step order inside a round
(warp w: top row w with bot row (w + t) mod B) and the once-per-sweep pairs inside a block.  The
numeric replay rotates with the kernel's formula and must reproduce numpy's singular values."""
import numpy as np
import pytest


def block_at(pos, R, nb):
    if pos == 0:
        return 0
    return 1 + ((pos - 1 - R) % (nb - 1))


def push(top, bot):
    """One end-of-round exchange exactly as the kernel's destinations are wired."""
    C = len(top)
    ntop, nbot = [None] * C, [None] * C
    for i in range(C):
        # top rows
        if i == 0:
            ntop[0] = top[0]
        elif i == C - 1:
            nbot[C - 1] = top[i]
        else:
            ntop[i + 1] = top[i]
        # bot rows
        if i == 0:
            ntop[1] = bot[0]
        else:
            nbot[i - 1] = bot[i]
    if C == 1:
        raise AssertionError("the kernel needs C >= 2")
    # rank 0 == C - 1 cannot happen; for C == 2 rank 1 is the last CTA: its top went to its own bot
    assert all(v is not None for v in ntop + nbot)
    return ntop, nbot


@pytest.mark.parametrize("C", [2, 4, 8, 16])
def test_tournament_meets_every_block_pair_once_and_ids_follow_the_closed_form(C):
    nb = 2 * C
    top = list(range(C))
    bot = [nb - 1 - i for i in range(C)]
    R = 0
    for sweep in range(3):
        met = set()
        for r in range(nb - 1):
            for i in range(C):
                assert top[i] == block_at(i, R, nb) and bot[i] == block_at(nb - 1 - i, R, nb)
                pair = frozenset((top[i], bot[i]))
                assert len(pair) == 2 and pair not in met
                met.add(pair)
            top, bot = push(top, bot)
            R += 1
        assert len(met) == nb * (nb - 1) // 2


def intra_pairs(B):
    """Pairs (p, q) of step t inside a block of B rows (B even), one warp pi < B/2 per pair."""
    out = []
    for t in range(B - 1):
        step = []
        for pi in range(B // 2):
            u = 0 if pi == 0 else ((pi - 1 + t) % (B - 1)) + 1
            v = ((B - 2 - pi + t) % (B - 1)) + 1
            step.append((min(u, v), max(u, v)))
        out.append(step)
    return out


@pytest.mark.parametrize("B", [2, 4, 6, 8, 16])
def test_intra_block_pairs_cover_the_block_once(B):
    seen = set()
    for step in intra_pairs(B):
        rows = [r for pq in step for r in pq]
        assert len(set(rows)) == len(rows) == B          # disjoint inside a step
        for pq in step:
            assert pq[0] != pq[1] and pq not in seen
            seen.add(pq)
    assert len(seen) == B * (B - 1) // 2


def rotate(x, y, kcols, tol):
    a = float(x[:kcols] @ x[:kcols]); b = float(y[:kcols] @ y[:kcols]); g = float(x[:kcols] @ y[:kcols])
    if g * g <= tol * tol * a * b or g == 0.0:
        return 0
    d, h = b - a, 2.0 * g
    r = np.sqrt(d * d + h * h)
    t = abs(h) / (abs(d) + r)
    if (d < 0) != (h < 0):
        t = -t
    c = 1.0 / np.sqrt(t * t + 1.0)
    s = c * t
    xo = x.copy()
    x[:] = c * xo - s * y
    y[:] = s * xo + c * y
    return 1


def replay(S, C, want_v, tol, max_sweeps=30):
    m, k = S.shape
    nb = 2 * C
    B = -(-m // nb)
    B += B & 1
    B = max(B, 2)
    L = k + (m if want_v else 0)
    rows = np.zeros((nb * B, L))
    rows[:m, :k] = S
    if want_v:
        rows[:m, k:] = np.eye(m)
    blocks = [rows[b * B:(b + 1) * B] for b in range(nb)]         # views
    top = list(range(C)); bot = [nb - 1 - i for i in range(C)]
    sweeps = 0
    for sweep in range(max_sweeps):
        rot = 0
        for r in range(nb - 1):
            for i in range(C):
                T, Bt = blocks[top[i]], blocks[bot[i]]
                if r == 0:
                    for step in intra_pairs(B):
                        half = B // 2
                        for w in range(B):
                            base = Bt if w // half else T
                            p, q = step[w % half]
                            rot += rotate(base[p], base[q], k, tol)
                for t in range(B):
                    for w in range(B):
                        rot += rotate(T[w], Bt[(w + t) % B], k, tol)
            top, bot = push(top, bot)
        sweeps += 1
        if rot == 0:
            break
    return rows[:m], sweeps


@pytest.mark.parametrize("m,k,C,want_v", [(24, 40, 2, True), (40, 64, 4, False), (37 + 1, 50, 4, True), (64, 64, 8, False)])
def test_numeric_replay_gives_the_singular_values(m, k, C, want_v):
    rng = np.random.RandomState(m + k)
    S = rng.standard_normal((m, k)) * np.logspace(0, -6, m)[:, None]
    tol = np.sqrt(k) * 1.1102230246251565e-16
    out, sweeps = replay(S, C, want_v, tol)
    s = np.sort(np.linalg.norm(out[:, :k], axis=1))[::-1]
    s_ref = np.linalg.svd(S, compute_uv=False)
    assert sweeps < 20
    assert np.max(np.abs(s - s_ref)) < 1e-13 * s_ref[0]
    if want_v:
        Vacc = out[:, k:]
        assert np.allclose(Vacc @ Vacc.T, np.eye(m), atol=1e-12)
        assert np.allclose(Vacc @ S, out[:, :k], atol=1e-12 * s_ref[0])


def test_numeric_replay_rank_deficient_block():
    """Numerically rank-deficient rows (singular values down to 1e-18 of the largest): noise-level
    rows keep rotating for a while, the iteration still ends and the values are numpy's."""
    rs = np.random.RandomState(5)
    m, k = 48, 64
    S = rs.standard_normal((m, m)) @ (np.logspace(0, -18, m)[:, None] * rs.standard_normal((m, k)))
    tol = np.sqrt(k) * 1.1102230246251565e-16
    s_ref = np.linalg.svd(S, compute_uv=False)
    out, sweeps = replay(S, 4, False, tol)
    s = np.sort(np.linalg.norm(out[:, :k], axis=1))[::-1]
    assert sweeps <= 30, sweeps
    assert np.max(np.abs(s - s_ref)) < 1e-13 * s_ref[0]
