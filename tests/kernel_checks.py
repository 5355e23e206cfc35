"""Per-kernel parity checks: every libvda operator against the plain PyTorch fp32 op it replaces, on the same
seeded tensors.  Each check returns (max_abs_err, tolerance).  Used by tests/test_kernels_gpu.py (asserting) and
by `python tests/kernel_checks.py` (prints a table and keeps going; handy for a first run on a new box)."""
from __future__ import annotations

import math
import os
import sys
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.nn.functional as F

from video_depth_anything_b200 import ops
from video_depth_anything_b200._lib import ACT_GELU, ACT_NONE, ACT_RELU, EPI_CONVT, EPI_GEGLU, EPI_LINEAR, EPI_TAIL

DEV = "cuda"
# the references are plain fp32 ops: no TF32 anywhere (cuDNN convolutions use it by default, which rounds the weights to
# 10 mantissa bits -- exactly what a 16-bit-weight kernel does, so a TF32 reference would hide weight-rounding error)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def _rand(shape, seed, scale=1.0, dtype=torch.float32):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(DEV).to(dtype)


def _err(a, ref):
    return (a.float() - ref.float()).abs().max().item()


def _tol(ref, dt, k=1):
    # one 16-bit rounding of the output (+ accumulated operand rounding ~ sqrt(k))
    eps = 2 ** -8 if dt == torch.bfloat16 else 2 ** -11
    return eps * (ref.float().abs().max().item() + 1e-3) * 2.0 + 1e-5


# ------------------------------------------------------------------------------------------------ GEMM
def check_gemm_plain(M, N, K, dt, bias=True, gamma=False, act=ACT_NONE, res1=None, res2=False, out_f32=False,
                     out_relu=False, seed=0, inplace=None):
    a = _rand((M, K), seed, 1.0, dt)
    w = _rand((N, K), seed + 1, 1.0 / math.sqrt(K), dt)
    b = _rand((N,), seed + 2) if bias else None
    g = (1.0 + _rand((N,), seed + 3, 0.1)) if gamma else None
    ref = a.float() @ w.float().t()
    if b is not None:
        ref = ref + b
    if g is not None:
        ref = ref * g
    if act == ACT_GELU:
        ref = F.gelu(ref)
    elif act == ACT_RELU:
        ref = F.relu(ref)
    r1 = None
    if res1 == "f32":
        r1 = _rand((M, N), seed + 4)
        ref = ref + r1
    elif res1 == "h16":
        r1 = _rand((M, N), seed + 4, 1.0, dt)
        ref = ref + r1.float()
    r2 = _rand((M, N), seed + 5, 1.0, dt) if res2 else None
    if r2 is not None:
        ref = ref + r2.float()
    inplace = (res1 == "f32" and out_f32) if inplace is None else inplace
    out = r1 if inplace else torch.empty(M, N, device=DEV, dtype=torch.float32 if out_f32 else dt)
    orl = torch.empty(M, N, device=DEV, dtype=dt) if out_relu else None
    ops.gemm(a, w, out, bias=b, gamma=g, act=act, res1=r1, res2=r2, out_relu=orl)
    torch.cuda.synchronize()
    e = _err(out, ref)
    if orl is not None:
        e = max(e, _err(orl, F.relu(ref)))
    return e, _tol(ref, dt)


def _stats_ref(x, parts):
    """(mean, M2) of every row over `parts` consecutive column groups, fp64."""
    M, C = x.shape
    g = x.double().reshape(M, parts, C // parts)
    mean = g.mean(-1)
    return torch.stack([mean, ((g - mean[..., None]) ** 2).sum(-1)], -1)


def check_gemm_residual_fold_producer(M, N, K, dt, seed=0, offset=0.0):
    """proj / fc2 epilogue with the LayerNorm-fold outputs (gemm.cu SPEC 4: residual box in and result box out by TMA):
    out = (a W^T + b) * gamma + out in place (fp32), out16 = h16(out), row statistics = per-row (mean, M2) over the
    kernel's column groups.  `offset`: residual rows with |mean| >> std (the shifted sums must keep M2)."""
    a = _rand((M, K), seed, 1.0, dt)
    w = _rand((N, K), seed + 1, 1.0 / math.sqrt(K), dt)
    b = _rand((N,), seed + 2)
    g = 1.0 + _rand((N,), seed + 3, 0.1)
    res = _rand((M, N), seed + 4) + offset
    ref = (a.float() @ w.float().t() + b) * g + res
    parts, part_cols = ops.rowstat_layout(M, N)
    assert parts * part_cols == N
    out = res.clone()
    o16 = torch.zeros(M, N, device=DEV, dtype=dt)
    stats = torch.zeros(M, parts, 2, device=DEV)
    ops.gemm(a, w, out, bias=b, gamma=g, res1=out, out16=o16, row_stats_out=stats)
    torch.cuda.synchronize()
    e = _err(out, ref)
    assert torch.equal(o16, out.to(dt)), "out16 is not the 16-bit rounding of the fp32 rows"
    sref = _stats_ref(out, parts)
    e_mean = (stats[..., 0].double() - sref[..., 0]).abs().max().item()
    rel_m2 = ((stats[..., 1].double() - sref[..., 1]).abs() / (sref[..., 1] + 1e-6)).max().item()
    assert e_mean < 1e-4 * (1 + abs(offset)) and rel_m2 < 2e-3, f"row statistics off: mean {e_mean:.3e} M2 rel {rel_m2:.3e}"
    return e, _tol(ref - res, dt) + 1e-6 * (1 + abs(offset))


def check_gemm_ln_fold(M, N, K, dt, act=ACT_NONE, seed=0, offset=0.3):
    """qkv / fc1 with the preceding LayerNorm folded into the epilogue (gemm.cu SPEC 5 / 6) against
    F.layer_norm(x) W^T + b (+ GELU) in fp32; x16 + statistics come from rowstats_cast (the chain's first link)."""
    x = _rand((M, K), seed, 1.5) + offset
    lw = 1.0 + _rand((K,), seed + 1, 0.2)
    lb = _rand((K,), seed + 2, 0.2)
    W = _rand((N, K), seed + 3, 1.0 / math.sqrt(K))
    b = _rand((N,), seed + 4)
    ref = F.layer_norm(x, (K,), lw, lb, 1e-6) @ W.t() + b
    if act == ACT_GELU:
        ref = F.gelu(ref)
    parts, part_cols = ops.rowstat_layout(M, K)
    x16 = torch.empty(M, K, device=DEV, dtype=dt)
    stats = torch.empty(M, parts, 2, device=DEV)
    ops.rowstats_cast(x, x16, stats)
    torch.cuda.synchronize()
    assert torch.equal(x16, x.to(dt))
    sref = _stats_ref(x, parts)
    assert (stats[..., 0].double() - sref[..., 0]).abs().max().item() < 1e-4
    assert ((stats[..., 1].double() - sref[..., 1]).abs() / (sref[..., 1] + 1e-6)).max().item() < 1e-3
    wf = (W * lw[None, :]).to(dt).contiguous()
    c1 = wf.float().sum(1).contiguous()
    c2 = (W @ lb + b).contiguous()
    out = torch.empty(M, N, device=DEV, dtype=dt)
    ops.gemm(x16, wf, out, bias=c2, act=act, ln_fold=(stats, c1, 1e-6))
    torch.cuda.synchronize()
    # operand rounding of the un-normalised rows: one 16-bit rounding of x relative to its own magnitude, seen through
    # rstd ~ 1/std(x); the plain path rounds LN(x) instead -- same order when |mean| is not far above std
    return _err(out, ref), _tol(ref, dt) * (2.0 + abs(offset))


def check_gemm_split(M, N, K, dt, conv=None, seed=0):
    """Validation precision: weights as a hi | lo pair of 16-bit matrices (ops.split_hi_lo, vda_gemm_params.a_k), A walked
    twice.  Against the fp32 product with the UNROUNDED weights: only A's rounding is left, so the error must be well
    below the plain 16-bit-weight GEMM's on the same data.  conv=(n,H,W,C): the implicit 3x3 conv walks its taps twice."""
    g = torch.Generator().manual_seed(seed)
    if conv is None:
        a = _rand((M, K), seed, 1.0, dt)
        w32 = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(DEV)
        ref = a.float() @ w32.t()
        out_s = torch.empty(M, N, device=DEV, dtype=torch.float32)
        out_p = torch.empty(M, N, device=DEV, dtype=torch.float32)
        ops.gemm(a, ops.split_hi_lo(w32, dt), out_s)
        ops.gemm(a, w32.to(dt).contiguous(), out_p)
    else:
        n, H, W, Ci = conv
        from video_depth_anything_b200.engine import pack_conv3x3
        x = _rand((n, Ci, H, W), seed, 1.0, dt)
        w4 = (torch.randn(N, Ci, 3, 3, generator=g) / math.sqrt(9 * Ci)).to(DEV)
        ref = F.conv2d(x.float(), w4, None, padding=1).permute(0, 2, 3, 1).reshape(n * H * W, N)
        a = x.permute(0, 2, 3, 1).contiguous()
        wp = pack_conv3x3(w4, Ci, N)
        out_s = torch.empty(n * H * W, N, device=DEV, dtype=torch.float32)
        out_p = torch.empty(n * H * W, N, device=DEV, dtype=torch.float32)
        ops.gemm(a, ops.split_hi_lo(wp, dt), out_s, conv_shape=(n, H, W, Ci))
        ops.gemm(a, wp.to(dt).contiguous(), out_p, conv_shape=(n, H, W, Ci))
    torch.cuda.synchronize()
    e_s, e_p = _err(out_s, ref), _err(out_p, ref)
    rms_s = (out_s - ref).pow(2).mean().sqrt().item()
    rms_p = (out_p - ref).pow(2).mean().sqrt().item()
    assert rms_s < 0.85 * rms_p, f"hi|lo weights do not reduce the error: rms {rms_s:.3e} vs plain {rms_p:.3e}"
    return e_s, _tol(ref, dt)


def check_gemm_patch_rowmap(dt, frames=3, P=20, N=384, K=592, seed=0):
    """row_group remap used by the patch-embed GEMM: out row = m + m/P + 1, res1 row = m%P + 1."""
    M = frames * P
    a = _rand((M, K), seed, 1.0, dt)
    w = _rand((N, K), seed + 1, 1.0 / math.sqrt(K), dt)
    b = _rand((N,), seed + 2)
    pos = _rand((P + 1, N), seed + 3)
    out = torch.zeros(frames * (P + 1), N, device=DEV)
    ops.gemm(a, w, out, bias=b, res1=pos, row_group=P)
    torch.cuda.synchronize()
    ref = (a.float() @ w.float().t() + b).reshape(frames, P, N) + pos[1:]
    o = out.reshape(frames, P + 1, N)
    e = max(_err(o[:, 1:], ref), o[:, 0].abs().max().item())
    return e, _tol(ref, dt)


def check_gemm_geglu(M, C, dt, seed=0):
    inner = 4 * C
    half = 128 if inner % 128 == 0 else 64
    a = _rand((M, C), seed, 1.0, dt)
    w = _rand((2 * inner, C), seed + 1, 1.0 / math.sqrt(C), dt)
    b = _rand((2 * inner,), seed + 2, 0.1)
    proj = a.float() @ w.float().t() + b
    ref = proj[:, :inner] * F.gelu(proj[:, inner:])
    from video_depth_anything_b200.engine import pack_geglu
    wp, bp = pack_geglu(w, b, half)
    out = torch.empty(M, inner, device=DEV, dtype=dt)
    ops.gemm(a, wp, out, bias=bp, epilogue=EPI_GEGLU, geglu_half=half)
    torch.cuda.synchronize()
    return _err(out, ref), _tol(ref, dt)


def check_gemm_convt(n, h, w_, ci, co, s, dt, seed=0):
    x = _rand((n, ci, h, w_), seed, 1.0, dt)
    wt = _rand((ci, co, s, s), seed + 1, 1.0 / math.sqrt(ci), dt)
    b = _rand((co,), seed + 2)
    ref = F.conv_transpose2d(x.float(), wt.float(), b, stride=s).permute(0, 2, 3, 1)    # NHWC
    from video_depth_anything_b200.engine import pack_convt
    co_pad = (co + 7) // 8 * 8
    wp, bp = pack_convt(wt, b, co_pad)
    a = x.permute(0, 2, 3, 1).reshape(n * h * w_, ci).contiguous()
    out = torch.zeros(n * h * s * w_ * s, co_pad, device=DEV, dtype=dt)
    ops.gemm(a, wp, out, bias=bp, epilogue=EPI_CONVT, convt=(s, co_pad, h, w_))
    torch.cuda.synchronize()
    o = out.reshape(n, h * s, w_ * s, co_pad)[..., :co]
    return _err(o, ref), _tol(ref, dt)


def check_conv3x3(n, H, W, ci, co, dt, bias=True, act=ACT_NONE, res=False, out_relu=False, seed=0):
    x = _rand((n, ci, H, W), seed, 1.0, dt)
    wt = _rand((co, ci, 3, 3), seed + 1, 1.0 / math.sqrt(9 * ci), dt)
    b = _rand((co,), seed + 2) if bias else None
    ref = F.conv2d(x.float(), wt.float(), b, padding=1)
    if act == ACT_RELU:
        ref = F.relu(ref)
    ref = ref.permute(0, 2, 3, 1)
    from video_depth_anything_b200.engine import pack_conv3x3
    wp = pack_conv3x3(wt, ci, co)
    a = x.permute(0, 2, 3, 1).contiguous()
    r1 = r2 = None
    if res:
        r1 = _rand((n * H * W, co), seed + 3, 1.0, dt)
        r2 = _rand((n * H * W, co), seed + 4, 1.0, dt)
        ref = ref + r1.float().reshape(ref.shape) + r2.float().reshape(ref.shape)
    out = torch.empty(n * H * W, co, device=DEV, dtype=dt)
    orl = torch.empty_like(out) if out_relu else None
    ops.gemm(a, wp, out, bias=b, act=act, res1=r1, res2=r2, out_relu=orl, conv_shape=(n, H, W, ci))
    torch.cuda.synchronize()
    e = _err(out.reshape(ref.shape), ref)
    if orl is not None:
        e = max(e, _err(orl.reshape(ref.shape), F.relu(ref)))
    return e, _tol(ref, dt)


def check_conv_tail(n, H, W, ci, dt, seed=0):
    x = _rand((n, ci, H, W), seed, 1.0, dt)
    w1 = _rand((32, ci, 3, 3), seed + 1, 1.0 / math.sqrt(9 * ci), dt)
    b1 = _rand((32,), seed + 2, 0.1)
    w2 = _rand((32,), seed + 3, 0.2).abs()
    b2 = 0.05
    h = F.relu(F.conv2d(x.float(), w1.float(), b1, padding=1))
    ref = F.relu((h * w2.view(1, 32, 1, 1)).sum(1) + b2)
    from video_depth_anything_b200.engine import pack_conv3x3
    wp = pack_conv3x3(w1, ci, 32)
    a = x.permute(0, 2, 3, 1).contiguous()
    out = torch.empty(n * H * W, device=DEV, dtype=torch.float32)
    ops.gemm(a, wp, out, bias=b1, epilogue=EPI_TAIL, tail_w=w2, tail_b=b2, conv_shape=(n, H, W, ci))
    torch.cuda.synchronize()
    return _err(out.reshape(ref.shape), ref), _tol(ref, dt) * 4


def check_tail_fused(n, ih, iw, oh, ow, ci, dt, seed=0):
    """bilinear(align_corners) upsample + 3x3 conv + ReLU + 1x1 + ReLU in one kernel vs the fp32 torch ops (the
    reference rounds the upsampled map to 16 bit too: F.interpolate runs under autocast, dpt_temporal.py:94-100)."""
    x = _rand((n, ci, ih, iw), seed, 1.0, dt)
    w1 = _rand((32, ci, 3, 3), seed + 1, 1.0 / math.sqrt(9 * ci), dt)
    b1 = _rand((32,), seed + 2, 0.1)
    w2 = _rand((32,), seed + 3, 0.2).abs()
    b2 = 0.05
    up = F.interpolate(x.float(), size=(oh, ow), mode="bilinear", align_corners=True).to(dt).float()
    h = F.relu(F.conv2d(up, w1.float(), b1, padding=1))
    ref = F.relu((h * w2.view(1, 32, 1, 1)).sum(1) + b2)
    from video_depth_anything_b200.engine import pack_conv3x3
    wp = pack_conv3x3(w1, ci, 32)
    a = x.permute(0, 2, 3, 1).contiguous()
    out = torch.full((n, oh, ow), -1.0, device=DEV, dtype=torch.float32)
    ops.tail_fused(a, wp, b1, w2, b2, out, n, ih, iw, oh, ow, ci)
    torch.cuda.synchronize()
    return _err(out, ref), _tol(ref, dt) * 4


def check_preprocess(n, h0, w0, input_size, seed=0):
    """device preprocessing (uint8 -> /255 -> INTER_CUBIC resize -> normalise -> CHW) vs the oracle's restatement of
    OpenCV's generic algorithm (float rounding only) and vs stock cv2 as the reference calls it (util/transform.py;
    the IPP build of cv2 differs from OpenCV's own generic path by ~1e-4 of the [0,1] pixel range, i.e. ~4e-4 after
    the division by std)."""
    import numpy as np
    from oracle import vda_oracle as O
    rng = np.random.default_rng(seed)
    frames = rng.integers(0, 256, (n, h0, w0, 3), dtype=np.uint8)
    sel = [n - 1, 0, n // 2]
    ref = torch.from_numpy(np.stack([O.preprocess_frame_generic(frames[i], input_size) for i in sel]))
    stock = torch.from_numpy(np.stack([O.preprocess_frame(frames[i], input_size) for i in sel]))
    nh, nw = ref.shape[2:]
    out = ops.preprocess_frames(torch.from_numpy(frames).to(DEV), torch.tensor(sel, dtype=torch.int32, device=DEV), nh, nw)
    torch.cuda.synchronize()
    e_stock = _err(out, stock.to(DEV))
    assert e_stock < 1e-3, f"vs stock cv2: {e_stock}"
    return _err(out, ref.to(DEV)), 1e-5


def check_copy_frames(slots, n, elems, seed=0):
    """gather + scatter of per-frame slabs (feature cache) vs torch indexing; exact."""
    g = torch.Generator().manual_seed(seed)
    src = _rand((slots, elems), seed, 1.0, torch.bfloat16)
    si = torch.randint(0, slots, (n,), generator=g).to(torch.int32).to(DEV)
    di = torch.randperm(slots, generator=g)[:n].to(torch.int32).to(DEV)
    dst = torch.zeros(slots, elems, dtype=torch.bfloat16, device=DEV)
    ref = dst.clone()
    ref[di.long()] = src[si.long()]
    ops.copy_frames(src, si, dst, di, n)
    e1 = _err(dst, ref)
    out = torch.empty(n, elems, dtype=torch.bfloat16, device=DEV)
    ops.copy_frames(src, si, out, None, n)
    e2 = _err(out, src[si.long()])
    ops.copy_frames(out, None, dst, di, n)
    return max(e1, e2, _err(dst, ref)), 0.0


# ------------------------------------------------------------------------------------------------ others
def check_layernorm(rows, C, dt, in_f32=True, drop_group=0, pe=False, seed=0):
    x = _rand((rows, C), seed, 2.0) + 0.5
    if not in_f32:
        x = x.to(dt)
    w = 1 + _rand((C,), seed + 1, 0.1)
    b = _rand((C,), seed + 2, 0.1)
    ref = F.layer_norm(x.float(), (C,), w, b, 1e-6)
    pe_t = None
    rpf = 0
    if pe:
        frames, rpf = 4, rows // 4
        pe_t = _rand((frames, C), seed + 3)
        ref = ref + pe_t.repeat_interleave(rpf, dim=0)
    if drop_group:
        keep = torch.arange(rows, device=DEV) % drop_group != 0
        ref = ref[keep]
    out = torch.empty(ref.shape, device=DEV, dtype=dt)
    ops.layernorm(x, w, b, 1e-6, out, drop_group=drop_group, pe=pe_t, pe_rows_per_frame=rpf)
    torch.cuda.synchronize()
    return _err(out, ref), _tol(ref, dt)


def check_groupnorm(frames, hw, C, dt, seed=0):
    x = (_rand((frames, hw, C), seed, 1.5) + 0.3).to(dt)
    w = 1 + _rand((C,), seed + 1, 0.1)
    b = _rand((C,), seed + 2, 0.1)
    ref = F.group_norm(x.float().permute(0, 2, 1), 32, w, b, 1e-6).permute(0, 2, 1)
    out = torch.empty_like(x)
    ops.groupnorm(x, w, b, 1e-6, out, frames, hw)
    torch.cuda.synchronize()
    return _err(out, ref), _tol(ref, dt)


def check_attention_spatial(frames, N, heads, dt, seed=0):
    qkv = _rand((frames, N, 3, heads, 64), seed, 1.0, dt)
    q, k, v = qkv.float().permute(2, 0, 3, 1, 4)
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(frames, N, heads * 64)
    out = torch.empty(frames, N, heads * 64, device=DEV, dtype=dt)
    ops.attention_spatial(qkv, out, frames, N, heads)
    torch.cuda.synchronize()
    return _err(out, ref), _tol(ref, dt) * 2


def check_attention_spatial_growing(frames, N, heads, dt, mode, seed=0):
    """Inputs that force the online-softmax RESCALE branch (attention_spatial.cu: the TMEM round trip of O when the
    running row maximum grows by more than 2^8 in log2 units, i.e. by > 44.4 in raw q.k units):
      ramp     q.k grows linearly along the key index by ~70 per 128-key tile for even query rows and falls for odd
               rows (one warp holds rows that rescale at every tile next to rows that never do);
      outlier  moderate logits plus ONE key in the last (partial) key tile whose logit is ~ +200 above the rest;
      late     the row maximum sits in the last key for every row after a long flat stretch (l must survive the
               rescale of an accumulated sum).
    Reference: fp32 SDPA on the same 16-bit inputs."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    q = torch.randn(frames, N, heads, 64, generator=g) * 0.5
    k = torch.randn(frames, N, heads, 64, generator=g) * 0.5
    v = torch.randn(frames, N, heads, 64, generator=g)
    u = torch.ones(64) / 8.0                                   # unit vector
    pos = torch.linspace(-1.0, 1.0, N).view(1, N, 1, 1)
    if mode == "ramp":
        per_tile = 70.0                                        # raw-logit growth per 128 keys (> 44.4)
        amp = per_tile * N / 128.0 / 2.0                       # q.k = +-8 * amp/8 * pos ...
        sign = torch.where(torch.arange(N) % 2 == 0, 1.0, -1.0).view(1, N, 1, 1)
        q = q + sign * 8.0 * u
        k = k + pos * (amp / 8.0) * u
    elif mode == "outlier":
        q = q + 6.0 * u
        k[:, N - 1] += 36.0 * u                                # q.k ~ +216 for the very last key
    elif mode == "late":
        q = q + 6.0 * u
        k[:, N - 1] += 9.7 * u                               # e^(58/8) ~ N: both parts weigh about the same
        k[:, : N - 1] *= 0.05                                  # flat stretch: l accumulates ~N before the jump
    qkv = torch.stack([q, k, v], dim=2).to(DEV).to(dt).contiguous()      # [frames, N, 3, heads, 64]
    qf, kf, vf = qkv.float().permute(2, 0, 3, 1, 4)
    ref = F.scaled_dot_product_attention(qf, kf, vf).transpose(1, 2).reshape(frames, N, heads * 64)
    out = torch.empty(frames, N, heads * 64, device=DEV, dtype=dt)
    ops.attention_spatial(qkv, out, frames, N, heads)
    torch.cuda.synchronize()
    return _err(out, ref), _tol(ref, dt) * 2


def check_groupnorm_offset(frames, hw, C, dt, offset, seed=0):
    """|mean| >> std (what ReLU'd feature maps of real checkpoints look like): E[x^2] - mean^2 in fp32 would lose
    the variance; the shifted / Chan-merged statistics must not.  The input is quantised to 16 bit first, the fp32
    reference sees the same values."""
    x = (_rand((frames, hw, C), seed, 1.0) + offset).to(dt)
    w = 1 + _rand((C,), seed + 1, 0.1)
    b = _rand((C,), seed + 2, 0.1)
    ref = F.group_norm(x.double().permute(0, 2, 1), 32, w.double(), b.double(), 1e-6).permute(0, 2, 1).float()
    out = torch.empty_like(x)
    ops.groupnorm(x, w, b, 1e-6, out, frames, hw)
    torch.cuda.synchronize()
    # one output rounding only: statistics errors would show up far above it
    eps = 2 ** -8 if dt == torch.bfloat16 else 2 ** -11
    return _err(out, ref), eps * (ref.abs().max().item()) + 1e-5


def check_attention_temporal(T, hw, C, dt, seed=0):
    heads = 8
    qkv = _rand((T * hw, 3 * C), seed, 1.0, dt)
    x = qkv.float().reshape(T, hw, 3, heads, C // heads)
    q, k, v = (x[:, :, i].permute(1, 2, 0, 3) for i in range(3))           # [hw, heads, T, dh]
    ref = F.scaled_dot_product_attention(q, k, v).permute(2, 0, 1, 3).reshape(T * hw, C)
    out = torch.empty(T * hw, C, device=DEV, dtype=dt)
    ops.attention_temporal(qkv, out, T, hw, C, heads)
    torch.cuda.synchronize()
    return _err(out, ref), _tol(ref, dt) * 2


def check_patch_im2col(frames, H, W, dt, seed=0):
    x = _rand((frames, 3, H, W), seed)
    hp, wp = H // 14, W // 14
    ref = F.unfold(x, kernel_size=14, stride=14).transpose(1, 2).reshape(frames * hp * wp, 588)
    out = torch.empty(frames * hp * wp, 592, device=DEV, dtype=dt)
    ops.patch_im2col(x, out)
    torch.cuda.synchronize()
    return max(_err(out[:, :588], ref), out[:, 588:].float().abs().max().item()), _tol(ref, dt)


def check_pos_embed(hp, wp, D=64, seed=0):
    pos = _rand((1 + 37 * 37, D), seed)
    sys.path.insert(0, "")
    from oracle.vda_oracle import interpolate_pos_encoding
    ref = interpolate_pos_encoding(pos.cpu().unsqueeze(0), hp, wp)[0].to(DEV)
    out = ops.pos_embed_bicubic(pos, hp, wp)
    torch.cuda.synchronize()
    return _err(out, ref), 2e-5


def check_im2col_s2(n, H, W, C, dt, seed=0):
    x = _rand((n, H, W, C), seed, 1.0, dt)
    out = ops.im2col3x3_s2(x, n, H, W, C)
    torch.cuda.synchronize()
    u = F.unfold(x.float().permute(0, 3, 1, 2), kernel_size=3, stride=2, padding=1)      # [n, C*9, L]
    oh, ow = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    ref = u.reshape(n, C, 9, oh * ow).permute(0, 3, 2, 1).reshape(n * oh * ow, 9 * C)
    return _err(out, ref), 0.0


def check_bilinear_nhwc(n, ih, iw, oh, ow, C, dt, seed=0):
    x = _rand((n, ih, iw, C), seed, 1.0, dt)
    ref = F.interpolate(x.float().permute(0, 3, 1, 2), size=(oh, ow), mode="bilinear", align_corners=True).permute(0, 2, 3, 1)
    out = torch.empty(n, oh, ow, C, device=DEV, dtype=dt)
    ops.bilinear_nhwc(x, out, n, ih, iw, oh, ow, C)
    torch.cuda.synchronize()
    return _err(out, ref), _tol(ref, dt)


def check_bilinear_f32(n, ih, iw, oh, ow, seed=0):
    x = _rand((n, ih, iw), seed)
    ref = F.interpolate(x.unsqueeze(1), size=(oh, ow), mode="bilinear", align_corners=True).squeeze(1)
    out = ops.bilinear_f32(x, oh, ow)
    torch.cuda.synchronize()
    return _err(out, ref), 1e-5


def check_add(n, dt, seed=0):
    a, b = _rand((n,), seed, 1.0, dt), _rand((n,), seed + 1, 1.0, dt)
    out = torch.empty_like(a)
    ops.add_h16(a, b, out)
    torch.cuda.synchronize()
    ref = a.float() + b.float()
    return _err(out, ref), _tol(ref, dt)


def check_lsq_affine(hw, seed=0):
    import numpy as np
    from oracle.vda_oracle import compute_scale_and_shift
    pred = _rand((2, hw), seed).abs() + 0.1
    target = pred * 1.7 + 0.25 + _rand((2, hw), seed + 1, 0.01)
    ss = torch.zeros(2, device=DEV)
    scratch = torch.zeros(4 * ops.LSQ_MAX_PARTIALS, device=DEV, dtype=torch.float64)
    ops.lsq_scale_shift(pred, target, ss, scratch)
    s_ref, t_ref = compute_scale_and_shift(pred.cpu().numpy().reshape(-1), target.cpu().numpy().reshape(-1))
    x = _rand((8, hw), seed + 2)
    prev = _rand((8, hw), seed + 3)
    w = torch.linspace(0, 1, 8, device=DEV)
    out = torch.empty_like(x)
    ops.affine_clamp_blend(x, ss, out, prev=prev, blend_w=w)
    out2 = torch.empty_like(x)
    ops.affine_clamp_blend(x, ss, out2)
    torch.cuda.synchronize()
    post = (x * float(s_ref) + float(t_ref)).clamp_min(0)
    ref = prev * (1 - w[:, None]) + post * w[:, None]
    e = max(abs(ss[0].item() - float(s_ref)), abs(ss[1].item() - float(t_ref)), _err(out, ref), _err(out2, post))
    return e, 2e-4


def check_align_chain(K, h, w, affine=True, seed=0):
    """The fused (scale, shift) recurrence (one cooperative kernel) against the per-window kernels WindowAligner.push
    launches (lsq_scale_shift + re-alignment of the key frame): the tables must be BIT-identical (1 GPU and N GPUs
    produce the same video), and close to the oracle's numpy restatement of utils/util.py:40-62."""
    import numpy as np
    from oracle.vda_oracle import compute_scale_and_shift
    g = torch.Generator().manual_seed(seed)
    base = torch.rand(1, 3, h, w, generator=g) + 0.2
    anchors = (base * (1.0 + 0.3 * torch.rand(K, 1, 1, 1, generator=g)) + 0.2 * torch.randn(K, 3, h, w, generator=g)
               - 0.1 * torch.arange(K).view(K, 1, 1, 1) / K).to(DEV).contiguous()
    scratch = torch.zeros(8 * ops.LSQ_MAX_PARTIALS, device=DEV, dtype=torch.float64)
    table = torch.full((K, 2), -7.0, device=DEV)
    ops.align_chain(anchors, table, scratch, affine=affine)
    # per-window path, exactly as WindowAligner.push sequences it
    ref_t = torch.zeros(K, 2, device=DEV)
    ref_t[0, 0] = 1.0
    ss = torch.tensor([1.0, 0.0], device=DEV)
    ref = torch.stack([anchors[0, 0], anchors[0, 2]])
    sc2 = torch.zeros(4 * ops.LSQ_MAX_PARTIALS, device=DEV, dtype=torch.float64)
    for k in range(1, K):
        if affine:                     # (.clone(): 16-byte aligned like WindowAligner's own buffers, whatever hw is)
            ops.lsq_scale_shift(anchors[k, 0:2].clone(), ref, ss, sc2)
        ref_t[k].copy_(ss)
        ops.affine_clamp_blend(anchors[k, 2:3].clone(), ss, ref[1:2])
    torch.cuda.synchronize()
    assert torch.equal(table, ref_t), f"chain table differs from the per-window kernels: {(table - ref_t).abs().max().item():.3e}"
    # oracle (numpy float32 sums): same recurrence on the host
    a = anchors.cpu().numpy()
    r0, r1 = a[0, 0], a[0, 2].copy()
    e = 0.0
    for k in range(1, K):
        s_, t_ = (1.0, 0.0)
        if affine:
            s_, t_ = compute_scale_and_shift(np.concatenate([a[k, 0].ravel(), a[k, 1].ravel()]), np.concatenate([r0.ravel(), r1.ravel()]))
        e = max(e, abs(float(s_) - table[k, 0].item()), abs(float(t_) - table[k, 1].item()))
        r1 = np.maximum(a[k, 2] * np.float32(table[k, 0].item()) + np.float32(table[k, 1].item()), 0)
    return e, 2e-4


BF, HF = torch.bfloat16, torch.float16
CHECKS = [
    # --- plain GEMM: tile / tail coverage, epilogues ---
    ("gemm 128x256x64 bf16", lambda: check_gemm_plain(128, 256, 64, BF)),
    ("gemm 128x256x64 fp16", lambda: check_gemm_plain(128, 256, 64, HF)),
    ("gemm 300x256x128 bf16 (M tail)", lambda: check_gemm_plain(300, 256, 128, BF)),
    ("gemm 1000x384x384 bf16 (bn=192)", lambda: check_gemm_plain(1000, 384, 384, BF)),
    ("gemm 777x48x384 bf16 (bn=48)", lambda: check_gemm_plain(777, 48, 384, BF)),
    ("gemm 2740x3072x1024 bf16 qkv-like", lambda: check_gemm_plain(2740, 3072, 1024, BF)),
    ("gemm 2740x1024x4096 bf16 ls+res f32 inplace", lambda: check_gemm_plain(2740, 1024, 4096, BF, gamma=True, res1="f32", out_f32=True)),
    ("gemm 2740x4096x1024 fp16 gelu", lambda: check_gemm_plain(2740, 4096, 1024, HF, act=ACT_GELU)),
    ("gemm 20000x1152x384 bf16 many tiles", lambda: check_gemm_plain(20000, 1152, 384, BF)),
    ("gemm 1369x256x256 bf16 res h16 x2 + relu copy", lambda: check_gemm_plain(1369, 256, 256, BF, res1="h16", res2=True, out_relu=True)),
    ("gemm 2740x1024x1024 bf16 residual + fold outputs (SPEC 4)", lambda: check_gemm_residual_fold_producer(2740, 1024, 1024, BF)),
    ("gemm 43840x1024x1024 bf16 residual + fold outputs", lambda: check_gemm_residual_fold_producer(43840, 1024, 1024, BF)),
    ("gemm 5000x1024x4096 fp16 residual + fold outputs", lambda: check_gemm_residual_fold_producer(5000, 1024, 4096, HF)),
    ("gemm 30140x384x1536 bf16 residual + fold outputs (bn=192)", lambda: check_gemm_residual_fold_producer(30140, 384, 1536, BF)),
    ("gemm 30140x384x384 fp16 residual + fold, mean=40", lambda: check_gemm_residual_fold_producer(30140, 384, 384, HF, offset=40.0)),
    ("gemm 2740x3072x1024 bf16 LN fold (SPEC 5)", lambda: check_gemm_ln_fold(2740, 3072, 1024, BF)),
    ("gemm 43840x4096x1024 bf16 LN fold + gelu (SPEC 6)", lambda: check_gemm_ln_fold(43840, 4096, 1024, BF, act=ACT_GELU)),
    ("gemm 30140x1152x384 fp16 LN fold", lambda: check_gemm_ln_fold(30140, 1152, 384, HF)),
    ("gemm 30140x1536x384 fp16 LN fold + gelu, mean=3", lambda: check_gemm_ln_fold(30140, 1536, 384, HF, act=ACT_GELU, offset=3.0)),
    ("gemm 2740x1024x4096 bf16 ls+res f32 separate out (SPEC 3)", lambda: check_gemm_plain(2740, 1024, 4096, BF, gamma=True, res1="f32", out_f32=True, inplace=False)),
    ("gemm hi|lo weights 2740x1024x1024 fp16", lambda: check_gemm_split(2740, 1024, 1024, HF)),
    ("gemm hi|lo weights 1000x384x592 fp16 (K tail)", lambda: check_gemm_split(1000, 384, 592, HF)),
    ("gemm hi|lo weights 3000x256x256 bf16", lambda: check_gemm_split(3000, 256, 256, BF)),
    ("conv3x3 hi|lo weights 2x37x37 256->256 fp16", lambda: check_gemm_split(0, 256, 0, HF, conv=(2, 37, 37, 256))),
    ("gemm K tail 592 bf16", lambda: check_gemm_plain(500, 384, 592, BF)),
    ("gemm K=24 bf16 (single partial k-block)", lambda: check_gemm_plain(500, 64, 24, BF)),
    ("gemm patch-embed row map bf16", lambda: check_gemm_patch_rowmap(BF)),
    ("gemm geglu C=256 bf16", lambda: check_gemm_geglu(1500, 256, BF)),
    ("gemm geglu C=192 fp16", lambda: check_gemm_geglu(700, 192, HF)),
    ("convT 4x4 s4 256ch bf16", lambda: check_gemm_convt(2, 5, 7, 256, 256, 4, BF)),
    ("convT 2x2 s2 96ch bf16", lambda: check_gemm_convt(2, 6, 5, 96, 96, 2, BF)),
    ("convT 4x4 s4 48ch fp16", lambda: check_gemm_convt(1, 6, 5, 48, 48, 4, HF)),
    # --- implicit-GEMM 3x3 conv ---
    ("conv3x3 2x37x37 256->256 bf16", lambda: check_conv3x3(2, 37, 37, 256, 256, BF)),
    ("conv3x3 1x19x33 1024->256 nobias bf16", lambda: check_conv3x3(1, 19, 33, 1024, 256, BF, bias=False)),
    ("conv3x3 1x74x74 64->64 fp16 relu", lambda: check_conv3x3(1, 74, 74, 64, 64, HF, act=ACT_RELU)),
    ("conv3x3 2x20x16 256->256 bf16 +2res +relu copy", lambda: check_conv3x3(2, 20, 16, 256, 256, BF, res=True, out_relu=True)),
    ("conv3x3 1x148x148 256->128 bf16", lambda: check_conv3x3(1, 148, 148, 256, 128, BF)),
    ("conv tail 1x70x84 128->32->1 bf16", lambda: check_conv_tail(1, 70, 84, 128, BF)),
    ("conv tail 1x56x70 64->32->1 fp16", lambda: check_conv_tail(1, 56, 70, 64, HF)),
    ("tail fused 2x40x48->70x84 128ch bf16", lambda: check_tail_fused(2, 40, 48, 70, 84, 128, BF)),
    ("tail fused 1x32x40->56x70 64ch fp16", lambda: check_tail_fused(1, 32, 40, 56, 70, 64, HF)),
    ("tail fused 1x9x11->14x14 128ch bf16 (one tile)", lambda: check_tail_fused(1, 9, 11, 14, 14, 128, BF)),
    ("tail fused 3x296x296->518x518 128ch bf16", lambda: check_tail_fused(3, 296, 296, 518, 518, 128, BF)),
    ("copy_frames 32 slots x 1369x1024", lambda: check_copy_frames(32, 22, 1369 * 1024)),
    ("copy_frames 7 slots x 8 elems", lambda: check_copy_frames(7, 5, 8)),
    ("preprocess 60x80 -> 98x126 (upscale)", lambda: check_preprocess(5, 60, 80, 98)),
    ("preprocess 720x1280 -> 518x924 (downscale)", lambda: check_preprocess(3, 720, 1280, 518)),
    ("preprocess 518x518 identity", lambda: check_preprocess(3, 518, 518, 518)),
    ("preprocess 300x1000 (aspect guard)", lambda: check_preprocess(3, 300, 1000, 518)),
    # --- norms ---
    ("layernorm 1370x1024 f32->bf16", lambda: check_layernorm(1370, 1024, BF)),
    ("layernorm 1370x384 f32->fp16 drop cls", lambda: check_layernorm(4 * 137, 384, HF, drop_group=137)),
    ("layernorm 400x192 h16 in + pe", lambda: check_layernorm(400, 192, BF, in_f32=False, pe=True)),
    ("layernorm 400x64 f32 + pe", lambda: check_layernorm(400, 64, BF, pe=True)),
    ("groupnorm 4x361x1024 bf16", lambda: check_groupnorm(4, 361, 1024, BF)),
    ("groupnorm 3x100x192 fp16", lambda: check_groupnorm(3, 100, 192, HF)),
    ("groupnorm 3x1369x64 bf16", lambda: check_groupnorm(3, 1369, 64, BF)),
    ("groupnorm 2x50x384 bf16", lambda: check_groupnorm(2, 50, 384, BF)),
    ("groupnorm 4x361x1024 bf16 mean=50 std", lambda: check_groupnorm_offset(4, 361, 1024, BF, 50.0)),
    ("groupnorm 2x5476x256 fp16 mean=30 std", lambda: check_groupnorm_offset(2, 5476, 256, HF, 30.0)),
    ("groupnorm 3x1369x64 fp16 mean=-200 std", lambda: check_groupnorm_offset(3, 1369, 64, HF, -200.0)),
    # --- attention ---
    ("attn spatial 2x1370x16 bf16", lambda: check_attention_spatial(2, 1370, 16, BF)),
    ("attn spatial 3x21x6 fp16", lambda: check_attention_spatial(3, 21, 6, HF)),
    ("attn spatial 1x2443x2 bf16", lambda: check_attention_spatial(1, 2443, 2, BF)),
    ("attn spatial 1x128x1 bf16", lambda: check_attention_spatial(1, 128, 1, BF)),
    ("attn spatial 4x300x3 bf16 (pair + split item)", lambda: check_attention_spatial(4, 300, 3, BF)),
    ("attn spatial 2x129x2 fp16 (2 tiles, partial 1)", lambda: check_attention_spatial(2, 129, 2, HF)),
    ("attn spatial 40x1370x16 bf16 (all CTAs, many items)", lambda: check_attention_spatial(40, 1370, 16, BF)),
    ("attn spatial 3x257x1 bf16 (split, n_kv=3)", lambda: check_attention_spatial(3, 257, 1, BF)),
    ("attn spatial 2x1370x6 fp16", lambda: check_attention_spatial(2, 1370, 6, HF)),
    # forced online-softmax rescale (never taken with unit-normal q/k)
    ("attn rescale ramp 2x1370x4 bf16", lambda: check_attention_spatial_growing(2, 1370, 4, BF, "ramp")),
    ("attn rescale ramp 2x1370x4 fp16", lambda: check_attention_spatial_growing(2, 1370, 4, HF, "ramp")),
    ("attn rescale ramp 1x2443x2 bf16", lambda: check_attention_spatial_growing(1, 2443, 2, BF, "ramp")),
    ("attn rescale ramp 1x2443x2 fp16", lambda: check_attention_spatial_growing(1, 2443, 2, HF, "ramp")),
    ("attn rescale ramp 3x129x2 bf16", lambda: check_attention_spatial_growing(3, 129, 2, BF, "ramp")),
    ("attn rescale ramp 3x129x2 fp16", lambda: check_attention_spatial_growing(3, 129, 2, HF, "ramp")),
    ("attn rescale outlier 2x1370x4 bf16", lambda: check_attention_spatial_growing(2, 1370, 4, BF, "outlier")),
    ("attn rescale outlier 2x1370x4 fp16", lambda: check_attention_spatial_growing(2, 1370, 4, HF, "outlier")),
    ("attn rescale outlier 1x2443x2 bf16", lambda: check_attention_spatial_growing(1, 2443, 2, BF, "outlier")),
    ("attn rescale outlier 3x129x2 fp16", lambda: check_attention_spatial_growing(3, 129, 2, HF, "outlier")),
    ("attn rescale late max 2x1370x4 bf16", lambda: check_attention_spatial_growing(2, 1370, 4, BF, "late")),
    ("attn rescale late max 1x2443x2 fp16", lambda: check_attention_spatial_growing(1, 2443, 2, HF, "late")),
    ("attn temporal 32x50x1024 bf16", lambda: check_attention_temporal(32, 50, 1024, BF)),
    ("attn temporal 32x77x256 fp16", lambda: check_attention_temporal(32, 77, 256, HF)),
    ("attn temporal 8x30x192 bf16", lambda: check_attention_temporal(8, 30, 192, BF)),
    ("attn temporal 32x30x64 bf16", lambda: check_attention_temporal(32, 30, 64, BF)),
    ("attn temporal 4x6x384 bf16", lambda: check_attention_temporal(4, 6, 384, BF)),
    # --- data movement ---
    ("patch im2col 2x56x70 bf16", lambda: check_patch_im2col(2, 56, 70, BF)),
    ("pos-embed bicubic 4x5", lambda: check_pos_embed(4, 5)),
    ("pos-embed bicubic 37x66", lambda: check_pos_embed(37, 66)),
    ("pos-embed bicubic 3x3", lambda: check_pos_embed(3, 3)),
    ("im2col s2 2x37x37x64 bf16", lambda: check_im2col_s2(2, 37, 37, 64, BF)),
    ("im2col s2 1x4x5x384 fp16", lambda: check_im2col_s2(1, 4, 5, 384, HF)),
    ("bilinear nhwc 19->37 bf16", lambda: check_bilinear_nhwc(2, 19, 19, 37, 37, 256, BF)),
    ("bilinear nhwc 296x300->518x525 fp16", lambda: check_bilinear_nhwc(1, 40, 44, 70, 77, 64, HF)),
    ("bilinear f32 70x84->60x80", lambda: check_bilinear_f32(3, 70, 84, 60, 80)),
    ("bilinear f32 identity", lambda: check_bilinear_f32(2, 56, 70, 56, 70)),
    ("add h16", lambda: check_add(8 * 1000, BF)),
    ("lsq + affine/clamp/blend", lambda: check_lsq_affine(60 * 80)),
    ("align chain K=12 60x80", lambda: check_align_chain(12, 60, 80)),
    ("align chain K=94 518x518", lambda: check_align_chain(94, 518, 518)),
    ("align chain K=5 15x14 (hw % 4 = 2)", lambda: check_align_chain(5, 15, 14)),
    ("align chain K=7 7x9 (hw odd)", lambda: check_align_chain(7, 7, 9)),
    ("align chain K=6 30x40 identity (metric)", lambda: check_align_chain(6, 30, 40, affine=False)),
    ("align chain K=1", lambda: check_align_chain(1, 30, 40)),
]


def main(filters=()):
    bad = 0
    for name, fn in CHECKS:
        inc = [f for f in filters if not f.startswith("-")]
        exc = [f[1:] for f in filters if f.startswith("-")]
        if (inc and not any(f in name for f in inc)) or any(f in name for f in exc):
            continue
        try:
            e, tol = fn()
            ok = e <= tol and e == e
            print(f"{'ok  ' if ok else 'FAIL'} {name:55s} err {e:.3e} tol {tol:.3e}", flush=True)
            bad += not ok
        except Exception as ex:  # noqa
            bad += 1
            print(f"EXC  {name:55s} {type(ex).__name__}: {ex}", flush=True)
            traceback.print_exc()
            if "CUDA" in str(ex) or "illegal" in str(ex):
                print("CUDA context is poisoned; stopping", flush=True)
                break
    print(f"{bad} failing checks", flush=True)
    return bad


if __name__ == "__main__":
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    sys.exit(1 if main(tuple(sys.argv[1:])) else 0)
