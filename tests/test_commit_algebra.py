"""Host model of blend_bwd.cu's commit round (CPU, numpy): the scalar (T, S) recurrence with pointer jumping over
same-pixel pairs against the four-channel, one-pair-after-the-other formulation of the reference's backward
(cuda_rasterizer/backward.cu:494-533: dL/dalpha = sum_ch (c_ch - accum_rec_ch) T dL/dpixel_ch - T_final/(1-alpha) bg.dL/dpixel,
walked back to front).

What it pins: (a) contracting "blended behind" with the pixel's upstream gradient BEFORE the recurrence gives the same
dL/dalpha, (b) folding the final-transmittance term into the start value of S is exact, (c) a pair is an affine map of (T, S),
maps compose as (m1, v1) then (m2, v2) = (m1 m2, v1 + m1 v2), and pointer jumping over the previous-peer links reproduces
the sequential result for every pair and for the state left behind — for any multiplicity pattern of a 32-lane round.
"""
import numpy as np
import pytest


def sequential_four_channel(pix, alpha, c, dp, T0, B0, T_final, tail):
    """One pair after the other in lane order (= list order, back to front); per-pixel state (T, B[4])."""
    T, B = T0.copy(), B0.copy()
    out = np.zeros(len(pix))
    Ti_out = np.zeros(len(pix))
    for i, p in enumerate(pix):
        rinv = 1.0 / (1.0 - alpha[i])
        Ti = T[p] * rinv                                   # transmittance in front of the entry
        out[i] = np.sum((c[i] * Ti - B[p] * rinv) * dp[p]) - T_final[p] * rinv * tail[p]
        B[p] = B[p] + c[i] * alpha[i] * Ti
        T[p] = Ti
        Ti_out[i] = Ti
    return out, Ti_out, T, B


def pointer_jumping_round(pix, alpha, c, dp, T0, S0):
    """The kernel's commit: lanes = pairs, prev = previous lane with the same pixel, log2 steps of composition."""
    n = len(pix)
    rinv = 1.0 / (1.0 - alpha)
    cdp = np.einsum("ic,ic->i", c, dp[pix])
    prev = np.full(n, -1)
    last_of = {}
    rank = np.zeros(n, dtype=int)
    for i, p in enumerate(pix):
        if p in last_of:
            prev[i] = last_of[p]
            rank[i] = rank[last_of[p]] + 1
        last_of[p] = i
    M, V = rinv.copy(), cdp * alpha * rinv
    span = int(rank.max()) if n else 0
    steps = 0
    while span > 0:
        Mp, Vp, pp = M[prev], V[prev], prev[prev]           # the shuffles: all lanes read before any lane updates
        live = prev >= 0
        V = np.where(live, Vp + Mp * V, V)
        M = np.where(live, Mp * M, M)
        prev = np.where(live, pp, prev)
        span >>= 1
        steps += 1
    Ti = T0[pix] * M
    Sa = S0[pix] + V * T0[pix]
    Sb = Sa - cdp * alpha * Ti
    dL_dalpha = cdp * Ti - Sb * rinv
    T1, S1 = T0.copy(), S0.copy()
    for p, i in last_of.items():                            # the chain's last pair writes the state back
        T1[p], S1[p] = Ti[i], Sa[i]
    return dL_dalpha, Ti, T1, S1, steps


def make_round(rng, n, n_pixels):
    pix = rng.integers(0, n_pixels, size=n)
    alpha = rng.uniform(1.0 / 255.0, 0.99, size=n)
    c = rng.uniform(0.0, 1.0, size=(n, 4))
    c[:, 3] = rng.uniform(0.5, 8.0, size=n)                 # depth channel
    dp = rng.normal(size=(32, 4))
    T0 = rng.uniform(1e-3, 1.0, size=32)
    B0 = rng.uniform(0.0, 1.0, size=(32, 4))
    T_final = T0 * rng.uniform(0.1, 1.0, size=32)
    tail = rng.normal(size=32)
    return pix, alpha, c, dp, T0, B0, T_final, tail


@pytest.mark.parametrize("n,n_pixels", [(32, 32), (32, 5), (32, 1), (17, 3), (1, 1), (32, 2)])
def test_scalar_recurrence_with_pointer_jumping_matches_sequential_four_channel(n, n_pixels):
    rng = np.random.default_rng(100 * n + n_pixels)
    for _ in range(20):
        pix, alpha, c, dp, T0, B0, T_final, tail = make_round(rng, n, n_pixels)
        want, Ti_w, T_w, B_w = sequential_four_channel(pix, alpha, c, dp, T0, B0, T_final, tail)
        S0 = np.einsum("pc,pc->p", B0, dp) + T_final * tail
        got, Ti_g, T_g, S_g, steps = pointer_jumping_round(pix, alpha, c, dp, T0, S0)
        longest = np.bincount(pix).max()
        assert steps == (0 if longest == 1 else int(np.ceil(np.log2(longest))))   # maxrank = longest - 1 halved to zero
        # the transmittance in front of the pixel's deepest pairs reaches T0 / (1 - alpha)^k: compare relatively
        np.testing.assert_allclose(Ti_g, Ti_w, rtol=1e-12)
        np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-9 * np.abs(want).max())
        np.testing.assert_allclose(T_g, T_w, rtol=1e-12)
        S_w = np.einsum("pc,pc->p", B_w, dp) + T_final * tail
        np.testing.assert_allclose(S_g, S_w, rtol=1e-9, atol=1e-9 * np.abs(S_w).max())


def test_rounds_chain_through_the_stored_state():
    """Two rounds back to back (the state a round leaves is the next round's start) = one sequential pass."""
    rng = np.random.default_rng(7)
    pix, alpha, c, dp, T0, B0, T_final, tail = make_round(rng, 64, 6)
    alpha = np.minimum(alpha, 0.6)                          # keep T finite over chains of ~10
    want, _, _, _ = sequential_four_channel(pix, alpha, c, dp, T0, B0, T_final, tail)
    S0 = np.einsum("pc,pc->p", B0, dp) + T_final * tail
    g1, _, T1, S1, _ = pointer_jumping_round(pix[:32], alpha[:32], c[:32], dp, T0, S0)
    g2, _, _, _, _ = pointer_jumping_round(pix[32:], alpha[32:], c[32:], dp, T1, S1)
    np.testing.assert_allclose(np.concatenate([g1, g2]), want, rtol=1e-9, atol=1e-9 * np.abs(want).max())
