"""CPU restatement of the synthetic PathTracker generator (pathtracker_models_b200/csrc/ptk_synth.cu).
TEST INFRASTRUCTURE ONLY.  Integer arithmetic throughout, so parity with the CUDA kernel is bit-exact.

The reference has no generator (its clips are external TFRecords, utils/TFRDataset.py:6-53); the
contract taken from it is the tensor layout/range (utils/engine.py:222-226) and "channel 2 of frame 0
is the goal map" (models/InT.py:236).  parity for this piece is therefore CUDA-vs-this-file only.
"""
import numpy as np

M64 = (1 << 64) - 1
DIRX = [65536, 60547, 46341, 25080, 0, -25080, -46341, -60547, -65536, -60547, -46341, -25080, 0, 25080, 46341, 60547]
DIRY = [0, 25080, 46341, 60547, 65536, 60547, 46341, 25080, 0, -25080, -46341, -60547, -65536, -60547, -46341, -25080]


def mix64(z):
    z = (z + 0x9E3779B97F4A7C15) & M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return z ^ (z >> 31)


def rnd32(key, k):
    return mix64((key + k * 0x632BE59BD9B4E019) & M64) >> 32


def synth(seed, first_index, B, T, H, W, D, speed):
    x = np.zeros((B, 3, T, H, W), np.float32)
    y = np.zeros((B,), np.float32)
    nd = max(D, 1) + 1
    lo, hix, hiy = 1 << 16, (W - 2) << 16, (H - 2) << 16
    for b in range(B):
        skey = mix64(seed ^ mix64((first_index + b) & M64))
        label = rnd32(skey, 0) & 1
        y[b] = label
        for dot in range(nd):
            key = mix64((skey + 0x1000003 * (dot + 1)) & M64)
            px = lo + rnd32(key, 1) % (hix - lo + 1)
            py = lo + rnd32(key, 2) % (hiy - lo + 1)
            d = rnd32(key, 3) & 15
            vx, vy = DIRX[d] * speed, DIRY[d] * speed
            curv = rnd32(key, 4) % 513 - 256
            drawn = dot == 0 or dot <= D
            sx, sy = (px + 32768) >> 16, (py + 32768) >> 16
            for t in range(T):
                if t > 0:
                    curv += rnd32(key, 8 + t) % 257 - 128
                    curv = max(-1024, min(1024, curv))
                    vx -= (vy * curv) >> 12      # python >> on negative ints is arithmetic, like the kernel
                    vy += (vx * curv) >> 12
                    px += vx
                    py += vy
                    if px < lo:
                        px, vx = 2 * lo - px, -vx
                    if px > hix:
                        px, vx = 2 * hix - px, -vx
                    if py < lo:
                        py, vy = 2 * lo - py, -vy
                    if py > hiy:
                        py, vy = 2 * hiy - py, -vy
                    px = max(lo, min(hix, px))
                    py = max(lo, min(hiy, py))
                if drawn:
                    x[b, 0, t, (py + 32768) >> 16, (px + 32768) >> 16] = 1
            fx, fy = (px + 32768) >> 16, (py + 32768) >> 16
            if dot == 0:
                x[b, 1, :, sy - 1:sy + 2, sx - 1:sx + 2] = 1
            if (dot == 0 and label == 1) or (dot == 1 and label == 0):
                x[b, 2, :, fy - 1:fy + 2, fx - 1:fx + 2] = 1
    return x, y
