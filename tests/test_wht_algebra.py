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


def half_reduce9(v):
    """k_blend_bwd_auto's reduction over the 16 lanes of each half-warp: v is (9, 32); returns (v0, v8) per lane."""
    v = [x.astype(np.float64).copy() for x in v]
    l16 = LANE & 15
    h = (l16 & 8) != 0
    for i in range(4):
        send, keep = np.where(h, v[i], v[i + 4]), np.where(h, v[i + 4], v[i])
        v[i] = keep + shfl_xor(send, 8)
    h = (l16 & 4) != 0
    for i in range(2):
        send, keep = np.where(h, v[i], v[i + 2]), np.where(h, v[i + 2], v[i])
        v[i] = keep + shfl_xor(send, 4)
    h = (l16 & 2) != 0
    send, keep = np.where(h, v[0], v[1]), np.where(h, v[1], v[0])
    v[0] = keep + shfl_xor(send, 2)
    v[0] = v[0] + shfl_xor(v[0], 1)
    for o in (8, 4, 2, 1):
        v[8] = v[8] + shfl_xor(v[8], o)
    return v[0], v[8]


def test_half_warp_reduction_and_lane_roles_of_the_autonomous_kernel():
    """Each half-warp reduces its own nine sums; the lane roles (target element + coefficient) of k_blend_bwd_auto turn
    them into exactly the gradient terms the default kernel's flush writes."""
    rng = np.random.default_rng(3)
    ddx, ddy = 960.0, 540.0
    # l16 -> (target, kA, kB, kC, kK, kO, takes_v8): the switch in the kernel
    roles = {0: ("mean_x", 2 * ddx, 0, 0, 0, 0, False), 1: ("mean_y", 0, ddy, 0, 0, 0, False),
             2: ("mean_x", 0, ddx, 0, 0, 0, False), 3: ("mean_y", 0, 0, 2 * ddy, 0, 0, False),
             4: ("conic_x", 0, 0, 0, -0.5, 0, False), 5: ("blue", 0, 0, 0, 1.0, 0, True),
             6: ("conic_y", 0, 0, 0, -1.0, 0, False), 8: ("conic_z", 0, 0, 0, -0.5, 0, False),
             10: ("opacity", 0, 0, 0, 0, 1.0, False), 12: ("red", 0, 0, 0, 1.0, 0, False),
             14: ("green", 0, 0, 0, 1.0, 0, False)}
    for _ in range(100):
        v = rng.normal(size=(9, 32))
        # every half-warp works on its own splat: a', b', c', opacity
        rec = rng.normal(size=(2, 4))
        rec[:, 3] = rng.uniform(0.05, 0.9, 2)
        v0, v8 = half_reduce9(v)
        for half in range(2):
            lanes = slice(16 * half, 16 * half + 16)
            s = v[:, lanes].sum(axis=1)
            ap, bp, cp, o = rec[half]
            want = {"mean_x": (2 * ap * s[0] + bp * s[1]) * ddx, "mean_y": (2 * cp * s[1] + bp * s[0]) * ddy,
                    "conic_x": -0.5 * s[2], "conic_y": -s[3], "conic_z": -0.5 * s[4], "opacity": s[5] / o,
                    "red": s[6], "green": s[7], "blue": s[8]}
            got = {k: 0.0 for k in want}
            for l16, (tgt, kA, kB, kC, kK, kO, use8) in roles.items():
                lane = 16 * half + l16
                coef = kA * ap + kB * bp + kC * cp + kK + kO / o
                got[tgt] += coef * (v8[lane] if use8 else v0[lane])
            for k in want:
                assert abs(got[k] - want[k]) <= 1e-9 * (1 + abs(want[k])), (k, got[k], want[k])


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
