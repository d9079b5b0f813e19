"""CPU transcription of spade_fused_fwd_v2_kernel (csrc/spade_fused.cu) for a whole small image: every per-lane address
computation of the kernel is followed through the ldmatrix / mma.sync.m16n8k16 fragment layouts (stages A1, A2, B, C, the
modulation, the image-border masks, ragged tiles) and the four outputs are compared with a plain PyTorch chain that rounds to
bf16 at the same points.  Checks the kernel's index logic without a GPU (the fragment layouts themselves were confirmed on B200
by version 1, which shares them)."""
import numpy as np
import torch
import torch.nn.functional as F

C, L, HID, TILE, XR, SR, AR, PSP, K3P = 64, 3, 4, 16, 22, 20, 18, 488, 88
N, H, W = 1, 20, 37
torch.manual_seed(0)
bf = lambda t: t.to(torch.bfloat16).float()
x = bf(torch.randn(N, C, H, W))
w1, b1 = torch.randn(L, C, 3, 3) * 0.1, torch.randn(L) * 0.3
w2, b2 = torch.randn(HID, L, 3, 3) * 0.4, torch.randn(HID) * 0.3
wg, bg = torch.randn(C, HID, 3, 3) * 0.3, torch.randn(C) * 0.3
wb, bb = torch.randn(C, HID, 3, 3) * 0.3, torch.randn(C) * 0.3

# ---- reference chain (bf16 storage of every intermediate, fp32 arithmetic) ----
seg_r = bf(F.conv2d(x, bf(w1), b1, padding=1))
act_r = bf(F.relu(F.conv2d(seg_r, bf(w2), b2, padding=1)))
gb_r = bf(F.conv2d(act_r, bf(torch.cat([wg, wb], 0)), torch.cat([bg, bb], 0), padding=1))
y_r = bf(x * (1 + gb_r[:, :C]) + gb_r[:, C:])

# ---- host operands exactly as ops._spade_fused_operands_build / the v2 branch build them ----
def pad_to(t, dim, size):
    if t.shape[dim] == size:
        return t
    shape = list(t.shape); shape[dim] = size - t.shape[dim]
    return torch.cat([t, t.new_zeros(shape)], dim)
w1m = bf(pad_to(w1.permute(2, 3, 0, 1).reshape(9 * L, C), 0, 32)).numpy()
p2 = bf(pad_to(pad_to(pad_to(w2.permute(0, 2, 3, 1), 3, 8).reshape(HID, 72), 1, 80), 0, 8)).numpy()
w3 = torch.cat([wg, wb], 0)
p3 = bf(pad_to(pad_to(w3.permute(0, 2, 3, 1), 3, 8).reshape(2 * C, 72), 1, 80)).numpy()
q1, q2, q3 = pad_to(b1, 0, 8).numpy(), pad_to(b2, 0, 8).numpy(), torch.cat([bg, bb]).numpy()
x_nhwc = x.permute(0, 2, 3, 1).contiguous().numpy()


def ldsm(addr_fn, nmat):
    regs = np.zeros((32, nmat, 2), np.float32)
    for i in range(nmat):
        rows = [addr_fn(8 * i + r) for r in range(8)]
        for lane in range(32):
            regs[lane, i] = rows[lane >> 2][2 * (lane & 3):2 * (lane & 3) + 2]
    return regs


def mma(acc, afrag, bfrag):
    A = np.zeros((16, 16), np.float32); B = np.zeros((16, 8), np.float32)
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        A[g, 2 * t:2 * t + 2] = afrag[lane, 0]; A[g + 8, 2 * t:2 * t + 2] = afrag[lane, 1]
        A[g, 2 * t + 8:2 * t + 10] = afrag[lane, 2]; A[g + 8, 2 * t + 8:2 * t + 10] = afrag[lane, 3]
        B[2 * t:2 * t + 2, g] = bfrag[lane, 0]; B[2 * t + 8:2 * t + 10, g] = bfrag[lane, 1]
    D = A @ B
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        acc[lane, 0:2] += D[g, 2 * t:2 * t + 2]; acc[lane, 2:4] += D[g + 8, 2 * t:2 * t + 2]


def r16(v):
    return torch.tensor(v, dtype=torch.float32).to(torch.bfloat16).float().numpy()


seg_o = np.zeros((N, H, W, 8), np.float32); act_o = np.zeros((N, H, W, 8), np.float32)
gb_o = np.zeros((N, H, W, 2 * C), np.float32); y_o = np.zeros((N, H, W, C), np.float32)
w2s = np.zeros((8, K3P), np.float32); w2s[:, :80] = p2
w3s = np.zeros((2 * C, K3P), np.float32); w3s[:, :80] = p3
tiles_x, tiles_y = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE
CH = C // 8
for tile in range(N * tiles_x * tiles_y):
    img, trem = divmod(tile, tiles_x * tiles_y)
    ty0, tx0 = (trem // tiles_x) * TILE, (trem % tiles_x) * TILE
    xs = np.zeros((XR * XR, CH, 8), np.float32)
    for q in range(XR * XR):
        iy, ix = ty0 - 3 + q // XR, tx0 - 3 + q % XR
        if 0 <= iy < H and 0 <= ix < W:
            for k in range(CH):
                xs[q, k ^ (q & 7)] = x_nhwc[img, iy, ix, 8 * k:8 * k + 8]
    segs = np.zeros((SR * SR + 1, 8), np.float32); acts = np.zeros((AR * AR + 1, 8), np.float32)
    ps = np.full((32, PSP), np.nan, np.float32)
    # ---- A1
    for warp in range(8):
        for nt in range(warp, (XR * XR + 7) // 8, 8):
            acc = [np.zeros((32, 4), np.float32) for _ in range(2)]
            for kc in range(C // 16):
                wa = np.zeros((2, 32, 4, 2), np.float32)
                for m in range(2):
                    for lane in range(32):
                        g, t = lane >> 2, lane & 3
                        r0, k0 = m * 16 + g, kc * 16 + 2 * t
                        wa[m, lane, 0] = w1m[r0, k0:k0 + 2]; wa[m, lane, 1] = w1m[r0 + 8, k0:k0 + 2]
                        wa[m, lane, 2] = w1m[r0, k0 + 8:k0 + 10]; wa[m, lane, 3] = w1m[r0 + 8, k0 + 8:k0 + 10]

                def baddr(lane, kc=kc, nt=nt):
                    bn, bhalf = lane & 7, (lane >> 3) & 1
                    q = min(nt * 8 + bn, XR * XR - 1)
                    return xs[q, (2 * kc + bhalf) ^ (q & 7)]
                b = ldsm(baddr, 2)
                for m in range(2):
                    mma(acc[m], wa[m], b)
            for lane in range(32):
                g, t = lane >> 2, lane & 3
                px = nt * 8 + 2 * t
                if px < PSP:
                    for m in range(2):
                        ps[m * 16 + g, px:px + 2] = acc[m][lane, 0:2]; ps[m * 16 + g + 8, px:px + 2] = acc[m][lane, 2:4]
    # ---- A2
    for pq in range(SR * SR):
        sy, sx = divmod(pq, SR)
        iy, ix = ty0 - 2 + sy, tx0 - 2 + sx
        inside = 0 <= iy < H and 0 <= ix < W
        v = np.zeros(8, np.float32)
        if inside:
            for c in range(L):
                a = np.float32(0)
                for tap in range(9):
                    a += ps[tap * L + c, (sy + tap // 3) * XR + sx + tap % 3]
                v[c] = a + q1[c]
        v = r16(v)
        segs[pq] = v
        if inside and 2 <= sy < 2 + TILE and 2 <= sx < 2 + TILE:
            seg_o[img, iy, ix] = v
    # ---- B
    for warp in range(8):
        for mt in range(warp, (AR * AR + 15) // 16, 8):
            acc = np.zeros((32, 4), np.float32)
            for j in range(5):
                def aaddr(lane, j=j, mt=mt):
                    arow, ahalf = lane & 15, lane >> 4
                    pa = mt * 16 + arow
                    aay, aax = divmod(pa, AR)
                    tap = 2 * j + ahalf
                    q = (aay + tap // 3) * SR + aax + tap % 3 if (pa < AR * AR and tap < 9) else SR * SR
                    return segs[q]

                def baddr(lane, j=j):
                    bn, bhalf = lane & 7, (lane >> 3) & 1
                    return w2s[bn, j * 16 + bhalf * 8: j * 16 + bhalf * 8 + 8]
                a4 = ldsm(aaddr, 4)
                mma(acc, a4, ldsm(baddr, 2))
            for lane in range(32):
                g, t = lane >> 2, lane & 3
                for hrow in range(2):
                    pq = mt * 16 + g + 8 * hrow
                    if pq >= AR * AR:
                        continue
                    ay, ax = divmod(pq, AR)
                    iy, ix = ty0 - 1 + ay, tx0 - 1 + ax
                    inside = 0 <= iy < H and 0 <= ix < W
                    v = r16(np.maximum(acc[lane, 2 * hrow:2 * hrow + 2] + q2[2 * t:2 * t + 2], 0)) if inside else np.zeros(2, np.float32)
                    acts[pq, 2 * t:2 * t + 2] = v
                    if inside and 1 <= ay < 1 + TILE and 1 <= ax < 1 + TILE:
                        act_o[img, iy, ix, 2 * t:2 * t + 2] = v
    # ---- C
    for warp in range(8):
        a2 = [[None] * 5 for _ in range(2)]
        for u in range(2):
            mt = warp + 8 * u
            for j in range(5):
                def aaddr(lane, j=j, mt=mt):
                    arow, ahalf = lane & 15, lane >> 4
                    pa = mt * 16 + arow
                    oay, oax = divmod(pa, TILE)
                    tap = 2 * j + ahalf
                    q = (oay + tap // 3) * AR + oax + tap % 3 if tap < 9 else AR * AR
                    return acts[q]
                a2[u][j] = ldsm(aaddr, 4)
        for cj in range(C // 8):
            ag = [np.zeros((32, 4), np.float32) for _ in range(2)]; ab = [np.zeros((32, 4), np.float32) for _ in range(2)]
            for j in range(5):
                def bg_addr(lane, j=j, cj=cj):
                    bn, bhalf = lane & 7, (lane >> 3) & 1
                    return w3s[cj * 8 + bn, j * 16 + bhalf * 8: j * 16 + bhalf * 8 + 8]

                def bb_addr(lane, j=j, cj=cj):
                    bn, bhalf = lane & 7, (lane >> 3) & 1
                    return w3s[C + cj * 8 + bn, j * 16 + bhalf * 8: j * 16 + bhalf * 8 + 8]
                bgm, bbt = ldsm(bg_addr, 2), ldsm(bb_addr, 2)
                for u in range(2):
                    mma(ag[u], a2[u][j], bgm); mma(ab[u], a2[u][j], bbt)
            for lane in range(32):
                g, t = lane >> 2, lane & 3
                ch = cj * 8 + 2 * t
                for u in range(2):
                    for hrow in range(2):
                        pq = (warp + 8 * u) * 16 + g + 8 * hrow
                        oy, ox = divmod(pq, TILE)
                        iy, ix = ty0 + oy, tx0 + ox
                        if not (iy < H and ix < W):
                            continue
                        gm = r16(ag[u][lane, 2 * hrow:2 * hrow + 2] + q3[ch:ch + 2])
                        bt = r16(ab[u][lane, 2 * hrow:2 * hrow + 2] + q3[C + ch:C + ch + 2])
                        q = (oy + 3) * XR + ox + 3
                        xv = xs[q, cj ^ (q & 7)][2 * t:2 * t + 2]
                        y_o[img, iy, ix, ch:ch + 2] = r16(xv * (1 + gm) + bt)
                        gb_o[img, iy, ix, ch:ch + 2] = gm; gb_o[img, iy, ix, C + ch:C + ch + 2] = bt


def rel(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).norm() / (b.norm() + 1e-30))


res = {"seg": rel(seg_o[..., :L], seg_r.permute(0, 2, 3, 1).numpy()), "actv": rel(act_o[..., :HID], act_r.permute(0, 2, 3, 1).numpy()),
       "gb": rel(gb_o, gb_r.permute(0, 2, 3, 1).numpy()), "y": rel(y_o, y_r.permute(0, 2, 3, 1).numpy())}
print(res)
assert np.abs(seg_o[..., L:]).max() == 0 and np.abs(act_o[..., HID:]).max() == 0
assert all(v < 5e-3 for v in res.values()), res
print("v2 kernel transcription matches the bf16 chain")
