#!/usr/bin/env python
"""Run every tensor-core layout probe in its own process (a faulting kernel poisons the CUDA context) and write one report.

    python scripts/tc_probe_matrix.py [--out gpurun_out/probes/matrix.json]
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PRELUDE = r'''
import sys, json, numpy as np, torch
sys.path.insert(0, %r)
from dist_tuto.pth_b200.ops import _ext
from dist_tuto.pth_b200.ops import tc_layouts as L
C = _ext.C(); dev = torch.device("cuda", 0)
def tma(t, dims, strides, box, sw, coords):
    img = C.tma_probe(torch.from_numpy(t.view(np.int16)).to(dev), dims, strides, box, sw, coords)
    torch.cuda.synchronize()
    return img.cpu().numpy()
def out(**kw):
    print("RESULT " + json.dumps(kw))
''' % ROOT

CASES = {
    # ---- TMA box loads (uint16 so values are exact)
    "tma2d_sw128_inner128": '''
t = np.arange(256 * 64, dtype=np.uint16).reshape(256, 64)
got = tma(t, [64, 256], [128], [64, 32], 3, [0, 8])
want = L.expected_tma_image(t[8:40])
out(ok=bool(np.array_equal(got, want)))
''',
    "tma4d_noswizzle_inner16": '''
B = 4
t = np.arange(B * 16 * 12 * 16, dtype=np.uint16).reshape(B, 16, 12, 16)
got = tma(t, [16, 12, 16, B], [32, 384, 6144], [8, 8, 16, 2], 0, [3, 2, 0, 1])
want = np.ascontiguousarray(t[1:3, :, 2:10, 3:11]).view(np.uint8).reshape(-1)
out(ok=bool(np.array_equal(got, want)), got=got[:32].tolist(), want=want[:32].tolist())
''',
    "tma4d_sw128_inner16": '''
B = 4
t = np.arange(B * 16 * 12 * 16, dtype=np.uint16).reshape(B, 16, 12, 16)
got = tma(t, [16, 12, 16, B], [32, 384, 6144], [8, 8, 16, 2], 3, [3, 2, 0, 1])
want = L.expected_tma_image(t[1:3, :, 2:10, 3:11])
out(ok=bool(np.array_equal(got, want)), got=got[:64].tolist(), want=want[:64].tolist())
''',
    "tma4d_sw128_inner16_origin": '''
B = 4
t = np.arange(B * 16 * 12 * 16, dtype=np.uint16).reshape(B, 16, 12, 16)
got = tma(t, [16, 12, 16, B], [32, 384, 6144], [8, 8, 16, 2], 3, [0, 0, 0, 0])
want = L.expected_tma_image(t[0:2, :, 0:8, 0:8])
out(ok=bool(np.array_equal(got, want)), got=got[:64].tolist(), want=want[:64].tolist())
''',
    "tma3d_sw128_inner16": '''
t = np.arange(16 * 12 * 16, dtype=np.uint16).reshape(16, 12, 16)
got = tma(t, [16, 12, 16], [32, 384], [8, 8, 16], 3, [3, 2, 0])
want = L.expected_tma_image(t[:, 2:10, 3:11])
out(ok=bool(np.array_equal(got, want)), got=got[:64].tolist(), want=want[:64].tolist())
''',
    "tma4d_sw128_inner128": '''
t = np.arange(4 * 8 * 16 * 64, dtype=np.uint16).reshape(4, 8, 16, 64)
got = tma(t, [64, 16, 8, 4], [128, 2048, 16384], [64, 4, 2, 2], 3, [0, 1, 2, 1])
want = L.expected_tma_image(t[1:3, 2:4, 1:5, :])
out(ok=bool(np.array_equal(got, want)))
''',
    "tma4d_sw32_inner16": '''
B = 4
t = np.arange(B * 16 * 12 * 16, dtype=np.uint16).reshape(B, 16, 12, 16)
got = tma(t, [16, 12, 16, B], [32, 384, 6144], [8, 8, 16, 2], 1, [3, 2, 0, 1])
out(ok=True, got=got[:128].tolist())
''',

    # ---- round 2, call 4: which small-inner-extent boxes does TMA accept?  (inner 16 B faulted with unaligned coords)
    "tma4d_inner16_aligned_x0": '''
B = 4
t = np.arange(B * 16 * 12 * 16, dtype=np.uint16).reshape(B, 16, 12, 16)
got = tma(t, [16, 12, 16, B], [32, 384, 6144], [8, 8, 16, 2], 0, [0, 2, 0, 1])
want = np.ascontiguousarray(t[1:3, :, 2:10, 0:8]).view(np.uint8).reshape(-1)
out(ok=bool(np.array_equal(got, want)))
''',
    "tma4d_inner16_aligned_x8": '''
B = 4
t = np.arange(B * 16 * 12 * 16, dtype=np.uint16).reshape(B, 16, 12, 16)
got = tma(t, [16, 12, 16, B], [32, 384, 6144], [8, 4, 16, 2], 0, [8, 2, 0, 1])
want = np.ascontiguousarray(t[1:3, :, 2:6, 8:16]).view(np.uint8).reshape(-1)
out(ok=bool(np.array_equal(got, want)))
''',
    "tma4d_nhwc_inner32_noswz": '''
B = 4
t = np.arange(B * 12 * 12 * 16, dtype=np.uint16).reshape(B, 12, 12, 16)
got = tma(t, [16, 12, 12, B], [32, 384, 4608], [16, 8, 8, 2], 0, [0, 3, 2, 1])
want = np.ascontiguousarray(t[1:3, 2:10, 3:11, :]).view(np.uint8).reshape(-1)
out(ok=bool(np.array_equal(got, want)))
''',
    "tma4d_nhwc_inner32_sw32": '''
B = 4
t = np.arange(B * 12 * 12 * 16, dtype=np.uint16).reshape(B, 12, 12, 16)
got = tma(t, [16, 12, 12, B], [32, 384, 4608], [16, 8, 8, 2], 1, [0, 3, 2, 1])
want = L.expected_tma_image_sw32(t[1:3, 2:10, 3:11, :])
out(ok=bool(np.array_equal(got, want)), got=got[:288:16].tolist(), want=want[:288:16].tolist())
''',
    "tma4d_nhwc_inner32_sw32_b1": '''
B = 5
t = np.arange(B * 12 * 12 * 16, dtype=np.uint16).reshape(B, 12, 12, 16)
got = tma(t, [16, 12, 12, B], [32, 384, 4608], [16, 8, 8, 1], 1, [0, 4, 4, 4])
want = L.expected_tma_image_sw32(t[4:5, 4:12, 4:12, :])
out(ok=bool(np.array_equal(got, want)))
''',
    "tma4d_nhwc_oob_batch_zero_fill": '''
B = 3
t = (np.arange(B * 12 * 12 * 16, dtype=np.uint16) + 1).reshape(B, 12, 12, 16)
got = tma(t, [16, 12, 12, B], [32, 384, 4608], [16, 8, 8, 2], 1, [0, 1, 1, 2])
box = np.zeros((2, 8, 8, 16), dtype=np.uint16); box[0] = t[2, 1:9, 1:9, :]
want = L.expected_tma_image_sw32(box)
out(ok=bool(np.array_equal(got, want)))
''',
    "umma_kmajor_sw32_a": '''
rng = np.random.default_rng(4)
a = L.bf16_bits(rng.standard_normal((128, 16)).astype(np.float32)); b = L.bf16_bits(rng.standard_normal((32, 64)).astype(np.float32))
ref = L.bits_to_f32(a) @ L.bits_to_f32(b)[:, :16].T
res = {}
for name, lbo, sbo in (("sbo256", 16, 256), ("sbo256_lbo256", 256, 256)):
    ops = [L.smem_desc(0, lbo, sbo, 6), L.smem_desc(0, 16, 1024, 2), 0, 0]
    d = C.umma_probe(torch.from_numpy(L.image_rows32(a)).to(dev), torch.from_numpy(L.image_rows128(b)).to(dev), L.idesc_bf16(128, 32), ops, 32)
    torch.cuda.synchronize()
    res[name] = float(np.abs(d.cpu().numpy() - ref).max() / np.abs(ref).max())
out(**res)
''',
    "umma_mn_major_sw32_a": '''
rng = np.random.default_rng(5); K = 64          # A[m = 128 = 8 taps x 16 ci][k = 64 positions], stored per tap as [k][16 m]
a = L.bf16_bits(rng.standard_normal((128, K)).astype(np.float32)); b = L.bf16_bits(rng.standard_normal((32, 64)).astype(np.float32))
blocks = [L.image_rows32(np.ascontiguousarray(a[i * 16:(i + 1) * 16].T)) for i in range(8)]
a_img = np.concatenate(blocks); atom = blocks[0].size            # 64 rows x 32 B = 2048
ref = L.bits_to_f32(a) @ L.bits_to_f32(b).T
res = {}
for name, lbo, sbo in (("cute", atom, 256), ("swapped", 256, atom)):
    ops = []
    for s in range(K // 16): ops += [L.smem_desc(s * 512, lbo, sbo, 6), L.smem_desc(s * 32, 16, 1024, 2), 0, int(s > 0)]
    d = C.umma_probe(torch.from_numpy(a_img).to(dev), torch.from_numpy(L.image_rows128(b)).to(dev), L.idesc_bf16(128, 32, a_mn=1), ops, 32)
    torch.cuda.synchronize()
    res[name] = float(np.abs(d.cpu().numpy() - ref).max() / np.abs(ref).max())
out(**res)
''',

    # ---- round 2, call 6: "shifted window" operands -- ONE image per tile in shared memory, the 25 taps are 25 descriptors
    "umma_window_fwd_kmajor_sw32": '''
rng = np.random.default_rng(6)
pix = L.bf16_bits(rng.standard_normal((12, 2, 12, 16)).astype(np.float32))          # [y][b][x][c]: 32-byte pixels
a_img = L.expected_tma_image_sw32(pix)
b = L.bf16_bits(rng.standard_normal((32, 64)).astype(np.float32)); b_img = L.image_rows128(b)
ky, kx = 2, 3
ops = [L.smem_desc(ky * 768 + kx * 32, 16, 384, 6), L.smem_desc(0, 16, 1024, 2), 0, 0]
d = C.umma_probe(torch.from_numpy(a_img).to(dev), torch.from_numpy(b_img).to(dev), L.idesc_bf16(128, 32), ops, 32).cpu().numpy()
torch.cuda.synchronize()
pf = L.bits_to_f32(pix); bf = L.bits_to_f32(b)[:, :16]
ref = np.zeros((128, 32), np.float32)
for oy in range(8):
    for bb in range(2):
        for ox in range(8):
            ref[(oy * 2 + bb) * 8 + ox] = pf[oy + ky, bb, ox + kx] @ bf.T
out(rel_err=float(np.abs(d - ref).max() / np.abs(ref).max()))
''',
    "umma_window_wgrad_mnmajor_sw32": '''
rng = np.random.default_rng(7)
pix = L.bf16_bits(rng.standard_normal((12, 12, 16)).astype(np.float32))             # one sample [y][x][c]
a_img = np.concatenate([L.expected_tma_image_sw32(pix), np.zeros(1024, np.uint8)])
b = L.bf16_bits(rng.standard_normal((32, 64)).astype(np.float32)); b_img = L.image_rows128(b)   # dC [co][pos]
ky = 1
ops = []
for ks in range(4): ops += [L.smem_desc(ky * 384 + ks * 768, 32, 384, 6), L.smem_desc(ks * 32, 16, 1024, 2), 0, int(ks > 0)]
d = C.umma_probe(torch.from_numpy(a_img).to(dev), torch.from_numpy(b_img).to(dev), L.idesc_bf16(128, 32, a_mn=1), ops, 32).cpu().numpy()
torch.cuda.synchronize()
pf = L.bits_to_f32(pix).reshape(144, 16); bf = L.bits_to_f32(b)
ref = np.zeros((128, 32), np.float32)
for kx in range(5):
    for ci in range(16):
        a_row = np.array([pf[(oy + ky) * 12 + ox + kx, ci] for oy in range(8) for ox in range(8)], np.float32)
        ref[kx * 16 + ci] = bf @ a_row
out(rel_err=float(np.abs(d[:80] - ref[:80]).max() / np.abs(ref[:80]).max()))
''',
    "tma4d_window_image_y_b_x_c": '''
B = 5
t = np.arange(B * 12 * 12 * 16, dtype=np.uint16).reshape(B, 12, 12, 16)
# dims (c, x, b, y): box {16, 12, 2, 12} -> shared image [y][b][x][c]
got = tma(t, [16, 12, B, 12], [32, 4608, 384], [16, 12, 2, 12], 1, [0, 0, 2, 0])
want = L.expected_tma_image_sw32(np.ascontiguousarray(t[2:4].transpose(1, 0, 2, 3)))
out(ok=bool(np.array_equal(got, want)))
''',
    # ---- UMMA descriptor modes
    "umma_kmajor": '''
rng = np.random.default_rng(0)
a = L.bf16_bits(rng.standard_normal((128, 64)).astype(np.float32)); b = L.bf16_bits(rng.standard_normal((32, 64)).astype(np.float32))
ops = []
for k in range(4): ops += [L.smem_desc(k * 32, 16, 1024), L.smem_desc(k * 32, 16, 1024), 0, int(k > 0)]
d = C.umma_probe(torch.from_numpy(L.image_rows128(a)).to(dev), torch.from_numpy(L.image_rows128(b)).to(dev), L.idesc_bf16(128, 32), ops, 32)
torch.cuda.synchronize()
ref = L.bits_to_f32(a) @ L.bits_to_f32(b).T
out(rel_err=float(np.abs(d.cpu().numpy() - ref).max() / np.abs(ref).max()))
''',
    "umma_mn_major_a": '''
rng = np.random.default_rng(1); K = 32
a = L.bf16_bits(rng.standard_normal((128, K)).astype(np.float32)); b = L.bf16_bits(rng.standard_normal((32, 64)).astype(np.float32))
blocks = [L.image_rows128(np.ascontiguousarray(a[i * 64:(i + 1) * 64].T)) for i in range(2)]
a_img = np.concatenate(blocks); atom = blocks[0].size
ref = L.bits_to_f32(a) @ L.bits_to_f32(b)[:, :K].T
res = {}
for name, lbo, sbo in (("cute", atom, 1024), ("swapped", 1024, atom)):
    ops = []
    for s in range(K // 16): ops += [L.smem_desc(s * 2048, lbo, sbo), L.smem_desc(s * 32, 16, 1024), 0, int(s > 0)]
    d = C.umma_probe(torch.from_numpy(a_img).to(dev), torch.from_numpy(L.image_rows128(b)).to(dev), L.idesc_bf16(128, 32, a_mn=1), ops, 32)
    torch.cuda.synchronize()
    res[name] = float(np.abs(d.cpu().numpy() - ref).max() / np.abs(ref).max())
out(**res)
''',
    "umma_wide_n": '''
rng = np.random.default_rng(2); K = 32
a = L.bf16_bits(rng.standard_normal((128, K)).astype(np.float32)); b = L.bf16_bits(rng.standard_normal((400, 64)).astype(np.float32))
blocks = [L.image_rows128(np.ascontiguousarray(a[i * 64:(i + 1) * 64].T)) for i in range(2)]
a_img = np.concatenate(blocks); atom = blocks[0].size
b_img = L.image_rows128(b)
ref = L.bits_to_f32(a) @ L.bits_to_f32(b)[:, :K].T
ops = []
for s in range(2): ops += [L.smem_desc(s * 2048, atom, 1024), L.smem_desc(s * 32, 16, 1024), 0, int(s > 0)]
d0 = C.umma_probe(torch.from_numpy(a_img).to(dev), torch.from_numpy(b_img).to(dev), L.idesc_bf16(128, 208, a_mn=1), ops, 208).cpu().numpy()
ops = []
for s in range(2): ops += [L.smem_desc(s * 2048, atom, 1024), L.smem_desc(208 * 128 + s * 32, 16, 1024), 0, int(s > 0)]
d1 = C.umma_probe(torch.from_numpy(a_img).to(dev), torch.from_numpy(b_img).to(dev), L.idesc_bf16(128, 192, a_mn=1), ops, 192).cpu().numpy()
torch.cuda.synchronize()
out(n208=float(np.abs(d0 - ref[:, :208]).max() / np.abs(ref).max()), n192=float(np.abs(d1 - ref[:, 208:]).max() / np.abs(ref).max()))
''',
    "umma_m64": '''
rng = np.random.default_rng(3)
a = L.bf16_bits(rng.standard_normal((64, 64)).astype(np.float32)); b = L.bf16_bits(rng.standard_normal((32, 64)).astype(np.float32))
ops = []
for k in range(4): ops += [L.smem_desc(k * 32, 16, 1024), L.smem_desc(k * 32, 16, 1024), 0, int(k > 0)]
d = C.umma_probe(torch.from_numpy(L.image_rows128(a)).to(dev), torch.from_numpy(L.image_rows128(b)).to(dev), L.idesc_bf16(64, 32), ops, 32).cpu().numpy()
torch.cuda.synchronize()
ref = L.bits_to_f32(a) @ L.bits_to_f32(b).T
lanes = np.array([32 * (r // 16) + (r % 16) for r in range(64)])
out(rel_err=float(np.abs(d[lanes] - ref).max() / np.abs(ref).max()))
''',
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "probes", "matrix.json"))
    ap.add_argument("--only", nargs="*")
    args = ap.parse_args()
    report = {}
    for name, body in CASES.items():
        if args.only and name not in args.only:
            continue
        env = dict(os.environ, CUDA_LAUNCH_BLOCKING="1")
        r = subprocess.run([sys.executable, "-c", PRELUDE + body], capture_output=True, text=True, timeout=300, env=env)
        res = None
        for line in r.stdout.splitlines():
            if line.startswith("RESULT "):
                res = json.loads(line[7:])
        tail = [ln for ln in r.stderr.strip().splitlines() if "Error" in ln or "error" in ln][-2:]
        report[name] = {"rc": r.returncode, "result": res, "stderr": tail}
        short = {k: (v if not isinstance(v, list) else "...") for k, v in (res or {}).items()}
        print(f"{name:32s} rc={r.returncode} {short} {tail[-1][:120] if tail and r.returncode else ''}", flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(report, f, indent=1)


if __name__ == "__main__":
    main()
