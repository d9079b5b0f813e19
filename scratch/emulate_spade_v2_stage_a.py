"""CPU emulation of stage A of spade_fused_fwd_v2_kernel (csrc/spade_fused.cu): follows the kernel's per-lane address arithmetic
through the documented ldmatrix / mma.sync.m16n8k16 fragment layouts and checks the resulting seg tile against a direct 3x3
convolution.  Verifies the index math (swizzle, fragment roles, P-buffer layout, shifted sum) without a GPU."""
import numpy as np

C, L, XR, SR, PSP = 64, 3, 22, 20, 488
rng = np.random.RandomState(0)
x_tile = rng.randn(XR * XR, C).astype(np.float32)              # xs pixel-major (already the zero-filled halo tile)
w1 = rng.randn(L, C, 3, 3).astype(np.float32)                   # x2map OIHW
b1 = rng.randn(L).astype(np.float32)

# host operand: [32][C], row = tap * L + c
w1m = np.zeros((32, C), np.float32)
w1m[:9 * L] = w1.transpose(2, 3, 0, 1).reshape(9 * L, C)

# ---- shared memory image of xs: 16-byte chunks (8 channels) swizzled by (pixel & 7)
CH = C // 8
xs = np.zeros((XR * XR, CH, 8), np.float32)
for q in range(XR * XR):
    for k in range(CH):
        xs[q, k ^ (q & 7)] = x_tile[q, 8 * k:8 * k + 8]


def ldsm_rows(addr_fn, nmat):
    """addr_fn(lane) -> the 8-element row that lane's address points at; returns per-lane register pairs for nmat matrices."""
    regs = np.zeros((32, nmat, 2), np.float32)
    for i in range(nmat):
        rows = [addr_fn(8 * i + r) for r in range(8)]           # matrix i: row r supplied by lane 8 i + r
        for lane in range(32):
            regs[lane, i] = rows[lane >> 2][2 * (lane & 3):2 * (lane & 3) + 2]
    return regs


def mma(afrag, bfrag):
    """afrag [32][4][2], bfrag [32][2][2] -> D fragment [32][4] (c0, c1 | c2, c3)."""
    A = np.zeros((16, 16), np.float32)
    B = np.zeros((16, 8), np.float32)
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        A[g, 2 * t:2 * t + 2] = afrag[lane, 0]; A[g + 8, 2 * t:2 * t + 2] = afrag[lane, 1]
        A[g, 2 * t + 8:2 * t + 10] = afrag[lane, 2]; A[g + 8, 2 * t + 8:2 * t + 10] = afrag[lane, 3]
        B[2 * t:2 * t + 2, g] = bfrag[lane, 0]; B[2 * t + 8:2 * t + 10, g] = bfrag[lane, 1]
    D = A @ B
    out = np.zeros((32, 4), np.float32)
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        out[lane, 0:2] = D[g, 2 * t:2 * t + 2]; out[lane, 2:4] = D[g + 8, 2 * t:2 * t + 2]
    return out


ps = np.full((32, PSP), np.nan, np.float32)
for warp in range(8):
    for nt in range(warp, (XR * XR + 7) // 8, 8):
        acc = np.zeros((2, 32, 4), np.float32)
        for kc in range(C // 16):
            # A fragments exactly as the kernel loads them from global memory
            wa = np.zeros((2, 32, 4, 2), np.float32)
            for m in range(2):
                for lane in range(32):
                    g, t = lane >> 2, lane & 3
                    r0, k0 = m * 16 + g, kc * 16 + 2 * t
                    wa[m, lane, 0] = w1m[r0, k0:k0 + 2]; wa[m, lane, 1] = w1m[r0 + 8, k0:k0 + 2]
                    wa[m, lane, 2] = w1m[r0, k0 + 8:k0 + 10]; wa[m, lane, 3] = w1m[r0 + 8, k0 + 8:k0 + 10]

            def baddr(lane):
                bn, bhalf = lane & 7, (lane >> 3) & 1
                q = min(nt * 8 + bn, XR * XR - 1)
                return xs[q, (2 * kc + bhalf) ^ (q & 7)]
            b = ldsm_rows(baddr, 2)
            for m in range(2):
                acc[m] += mma(wa[m], b)
        for lane in range(32):
            g, t = lane >> 2, lane & 3
            px = nt * 8 + 2 * t
            if px < PSP:
                for m in range(2):
                    ps[m * 16 + g, px:px + 2] = acc[m, lane, 0:2]
                    ps[m * 16 + g + 8, px:px + 2] = acc[m, lane, 2:4]

seg = np.zeros((SR * SR, L), np.float32)
for pq in range(SR * SR):
    sy, sx = divmod(pq, SR)
    for c in range(L):
        a = np.float32(0)
        for tap in range(9):
            a += ps[tap * L + c, (sy + tap // 3) * XR + sx + tap % 3]
        seg[pq, c] = a + b1[c]

# direct convolution on the same tile: seg region pixel (sy, sx) uses xs pixels (sy + r, sx + s)
xt = x_tile.reshape(XR, XR, C)
ref = np.zeros((SR, SR, L), np.float32)
for r in range(3):
    for s in range(3):
        ref += np.einsum("yxc,lc->yxl", xt[r:r + SR, s:s + SR], w1[:, :, r, s])
ref += b1
err = np.abs(seg.reshape(SR, SR, L) - ref).max() / np.abs(ref).max()
print("stage A (v2) emulation: max rel err %.2e" % err, "OK" if err < 1e-5 else "MISMATCH")
assert not np.isnan(seg).any() and err < 1e-5
