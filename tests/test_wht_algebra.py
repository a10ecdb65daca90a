"""CPU check of the algebra behind the EXPERIMENTAL backward reduction (csrc/blend.cu, k_blend_bwd_wht): a numpy
emulation of the 32-lane Walsh-Hadamard butterfly, of the 4-value colour butterfly and of the flush formulas, against
the directly summed moments.  The kernel itself is validated on the device (tests/test_gpu_parity.py with
GS_B200_EXPERIMENTAL=1); this pins the index maps and the collected coefficients the kernel hard-codes."""
import numpy as np

LANE = np.arange(32)


def shfl_xor(v, o):
    return v[LANE ^ o]


def slot_of(L):
    """lane -> storage slot of its coefficient (k_blend_bwd_wht: `slot`)."""
    pc = bin(L).count("1")
    if pc == 0:
        return 0
    if pc == 1:
        return (L & -L).bit_length()              # __ffs(lane) = 1 + bit index
    if pc == 2:
        i, j = (L & -L).bit_length() - 1, L.bit_length() - 1
        return 6 + (i * (9 - i)) // 2 + (j - i - 1)
    return -1


def warp_partial(m, c, F):
    """What one warp leaves in shared memory for one splat: 16 Walsh-Hadamard coefficients + 3 colour sums."""
    w = m.astype(F).copy()
    for s in range(5):
        sg = np.where((LANE >> s) & 1, F(-1), F(1)).astype(F)
        w = (sg * w + shfl_xor(w, 1 << s)).astype(F)
    cf = np.zeros(19, F)
    for L in LANE:
        if slot_of(int(L)) >= 0:
            cf[slot_of(int(L))] = w[L]
    h16, h8 = (LANE & 16) != 0, (LANE & 8) != 0
    c0, c1, c2 = (x.astype(F) for x in c)
    zero = np.zeros(32, F)
    r0 = (np.where(h16, c2, c0) + shfl_xor(np.where(h16, c0, c2), 16)).astype(F)
    r1 = (np.where(h16, zero, c1) + shfl_xor(np.where(h16, c1, zero), 16)).astype(F)
    q = (np.where(h8, r1, r0) + shfl_xor(np.where(h8, r0, r1), 8)).astype(F)
    for o in (4, 2, 1):
        q = (q + shfl_xor(q, o)).astype(F)
    for L in (0, 8, 16):
        cf[16 + (L >> 3)] = q[L]
    return cf


def flush(cf, ux, uy, F):
    """Flush formulas of the kernel: moments about the splat centre from one warp's coefficients."""
    W0, E0, E1, E2, E3, E4 = (F(x) for x in cf[:6])
    Lx = F(3.5) * W0 - F(0.5) * E0 - E1 - F(2) * E2
    Ly = F(1.5) * W0 - F(0.5) * E3 - E4
    Lxx = F(17.5) * W0 - F(3.5) * E0 - F(7) * E1 - F(14) * E2 + cf[6] + F(2) * cf[7] + F(4) * cf[10]
    Lyy = F(3.5) * W0 - F(1.5) * E3 - F(3) * E4 + cf[15]
    Lxy = F(0.25) * (F(21) * W0 - F(3) * E0 - F(6) * E1 - F(12) * E2 - F(7) * E3 - F(14) * E4 + cf[8] + F(2) * cf[9] +
                     F(2) * cf[11] + F(4) * cf[12] + F(4) * cf[13] + F(8) * cf[14])
    ux, uy = F(ux), F(uy)
    return np.array([ux * W0 - Lx, uy * W0 - Ly, ux * (ux * W0 - F(2) * Lx) + Lxx, ux * (uy * W0 - Ly) - uy * Lx + Lxy,
                     uy * (uy * W0 - F(2) * Ly) + Lyy, W0], dtype=np.float64)


def case(rng, far):
    m = rng.normal(size=32) * (rng.random(32) < 0.6)
    c = rng.normal(size=(3, 32))
    ax, ay = rng.uniform(-far, far, 2)
    warp = int(rng.integers(0, 8))
    X0, Y0 = 32.0, 48.0
    px = X0 + (warp & 1) * 8 + (LANE & 7)          # pixel_of_thread
    py = Y0 + (warp >> 1) * 4 + (LANE >> 3)
    ux, uy = ax - (X0 + (warp & 1) * 8), ay - (Y0 + (warp >> 1) * 4)
    return m, c, ax - px, ay - py, ux, uy


def test_slots_are_a_bijection_onto_16():
    slots = [slot_of(int(L)) for L in LANE]
    assert sorted(s for s in slots if s >= 0) == list(range(16))
    assert sum(1 for s in slots if s < 0) == 16


def test_coefficients_give_exact_moments_in_fp64():
    rng = np.random.default_rng(0)
    for _ in range(300):
        m, c, dx, dy, ux, uy = case(rng, 80.0)
        cf = warp_partial(m, c, np.float64)
        ref = np.array([(m * dx).sum(), (m * dy).sum(), (m * dx * dx).sum(), (m * dx * dy).sum(), (m * dy * dy).sum(), m.sum()])
        np.testing.assert_allclose(flush(cf, ux, uy, np.float64), ref, rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(cf[16:19], c.sum(axis=1), rtol=1e-12, atol=1e-12)


def test_fp32_error_is_of_the_order_of_direct_summation():
    """|error| / sum |m| |f| stays at a few fp32 ulps, also for splat centres hundreds of pixels from the block."""
    rng = np.random.default_rng(1)
    worst = 0.0
    for _ in range(1500):
        m, c, dx, dy, ux, uy = case(rng, float(rng.choice([5.0, 50.0, 400.0])))
        m32 = m.astype(np.float32).astype(np.float64)
        ref = np.array([(m32 * dx).sum(), (m32 * dy).sum(), (m32 * dx * dx).sum(), (m32 * dx * dy).sum(), (m32 * dy * dy).sum(),
                        m32.sum()])
        scale = np.array([(abs(m32) * abs(dx)).sum(), (abs(m32) * abs(dy)).sum(), (abs(m32) * dx * dx).sum(),
                          (abs(m32) * abs(dx * dy)).sum(), (abs(m32) * dy * dy).sum(), abs(m32).sum()]) + 1e-30
        got = flush(warp_partial(m, c, np.float32), ux, uy, np.float32)
        worst = max(worst, float((abs(got - ref) / scale).max()))
    assert worst < 2e-6, worst
