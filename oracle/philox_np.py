"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the sampler's on-device random draws.

The reference draws from numpy's global MT19937 stream; that stream is inherently sequential, so the device mode
uses a counter RNG instead (Philox4x32-10, Salmon et al. SC'11) and is compared with the reference
*distributionally*.  This file restates the device's draw construction (ogbench_b200/csrc/device_common.cuh:
key/counter layout, the 64-bit multiply-shift bounded integer, the 53-bit unit double, the geometric inversion)
so that tests can also check the device mode bit-for-bit: ``philox_draws(...)`` returns the same ``Draws`` record
the kernel uses internally, which is then fed to the oracle sampler.
"""

from __future__ import annotations

import numpy as np

from oracle.replay_oracle import Draws, GoalDraws

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint64(0x9E3779B9), np.uint64(0xBB67AE85)
MASK32 = np.uint64(0xFFFFFFFF)

PURPOSE_IDX, PURPOSE_GOAL, PURPOSE_GOAL_LOW, PURPOSE_CROP, PURPOSE_MIX, PURPOSE_TRL_MID, PURPOSE_COIN = 0, 1, 2, 3, 4, 6, 7


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10; all arguments uint64 arrays/scalars holding 32-bit values."""
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint64) for x in (c0, c1, c2, c3))
    k0, k1 = np.uint64(k0), np.uint64(k1)
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & MASK32, lo1, (hi0 ^ c3 ^ k1) & MASK32, lo0
        k0 = (k0 + W0) & MASK32
        k1 = (k1 + W1) & MASK32
    return c0, c1, c2, c3


def draw4(seed, stream, batch, rows, purpose):
    rows = np.asarray(rows, dtype=np.uint64)
    c1 = np.full_like(rows, np.uint64(batch & 0xFFFFFFFF))
    c2 = np.full_like(rows, np.uint64((batch >> 32) & 0xFFFFFFFF))
    c3 = np.full_like(rows, np.uint64(purpose | (stream << 8)))
    return philox4x32_10(rows, c1, c2, c3, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)


def bounded_u64(hi, lo, n):
    """floor(((hi << 32) | lo) * n / 2^64) -- exact, via Python integers."""
    x = (hi.astype(object) << 32) | lo.astype(object)
    return np.array([(int(v) * int(n)) >> 64 for v in x], dtype=np.int64)


def unit_double(a, b):
    return ((a >> np.uint64(5)).astype(np.float64) * 67108864.0 + (b >> np.uint64(6)).astype(np.float64)) * (1.0 / 9007199254740992.0)


def geometric_from_unit(u, discount):
    log_1mp = np.log(1.0 - (1.0 - discount))
    x = np.ceil(np.log(1.0 - u) / log_1mp)
    return np.maximum(x, 1.0).astype(np.int64)


def geometric_is_knife_edge(u, discount, tol=1e-9):
    """True where log(1-u)/log(1-p) lies within `tol` of an integer, i.e. where a 1-ulp difference between the
    device's and numpy's log could legitimately change the ceil."""
    log_1mp = np.log(1.0 - (1.0 - discount))
    q = np.log(1.0 - u) / log_1mp
    return np.abs(q - np.round(q)) < tol * np.maximum(1.0, np.abs(q))


def philox_draws(seed, stream, batch_index, batch_size, n_choices, goal_sets, aug, p_aug, padding=3, idxs_given=False):
    """Draws of one device sample() call.

    goal_sets: list of (slot, geom, discount, cur_only) with slot 0 = value, 1 = low-value, 2 = actor.  aug: whether
    the per-batch coin is drawn at all.  Word layout (device_common.cuh, enum Purpose): a goal set owns 64 random bits
    that serve either as its random-goal position or as its geometric / distance uniform -- the mix coins decide
    which one the sampler looks at, so building both from the same bits here is equivalent.
    """
    rows = np.arange(batch_size, dtype=np.uint64)
    d = Draws()
    w = draw4(seed, stream, batch_index, rows, PURPOSE_IDX)            # x,y: index position; z,w: value mix coins
    if not idxs_given:
        d.idx_pos = bounded_u64(w[0], w[1], n_choices)
    knife = np.zeros(batch_size, dtype=bool)
    gb = draw4(seed, stream, batch_index, rows, PURPOSE_GOAL)          # x,y: value goal bits; z,w: actor goal bits
    low = draw4(seed, stream, batch_index, rows, PURPOSE_GOAL_LOW)     # x,y: low-value goal bits; z,w: its mix coins
    amix = draw4(seed, stream, batch_index, rows, PURPOSE_MIX)         # x,y: actor mix coins
    two32 = 4294967296.0
    bits_of = {0: (gb[0], gb[1]), 1: (low[0], low[1]), 2: (gb[2], gb[3])}
    coins_of = {0: (w[2], w[3]), 1: (low[2], low[3]), 2: (amix[0], amix[1])}
    for slot, geom, discount, cur_only in goal_sets:
        a0, a1 = bits_of[slot]
        g = GoalDraws(rand_pos=bounded_u64(a0, a1, n_choices))
        u = unit_double(a0, a1)
        if geom:
            g.offset = geometric_from_unit(u, discount)
            knife |= geometric_is_knife_edge(u, discount)
        else:
            g.dist = u
        if not cur_only:
            # when the actor mix is not a real choice the device skips PURPOSE_MIX; the coins then cannot change the
            # outcome, so the values used here are immaterial
            c0, c1 = coins_of[slot]
            g.u_traj = c0.astype(np.float64) / two32
            g.u_cur = c1.astype(np.float64) / two32
        d.goals.append(g)
    if aug:
        c = draw4(seed, stream, batch_index, np.array([0xFFFFFFFF], dtype=np.uint64), PURPOSE_COIN)
        d.aug_coin = float(unit_double(c[0], c[1])[0])
        if d.aug_coin < p_aug:
            span = np.uint64(2 * padding + 1)
            cw = draw4(seed, stream, batch_index, rows, PURPOSE_CROP)
            joint = (cw[0] * span * span) >> np.uint64(32)                # (cy, cx) jointly uniform on span x span
            d.crop = np.stack([joint // span, joint % span], axis=1).astype(np.int64)
    return d, knife
