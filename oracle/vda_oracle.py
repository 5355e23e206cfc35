"""ORACLE — test infrastructure only.  Never imported by the product path.

A plain-PyTorch fp32 restatement of the reference's hot path
(`VideoDepthAnything.forward` / `infer_video_depth`) written as pure functions of
a state dict, so it runs where `/root/reference` does not exist (the GPU box).
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this module.

Parity pin: `oracle/make_golden.py` imports the real reference in the build
container, loads the same synthetic state dict (`strict=True`), and checks that
this restatement reproduces the reference's `forward` and `infer_video_depth`
(see tests/golden/MANIFEST.json for the measured max-abs differences); the
reference's own outputs are committed under tests/golden/ and re-checked by
`tests/test_oracle_golden.py`.  The reference ships no tests / golden vectors of
its own for this path.

Every function cites the reference file:line it restates (paths relative to
/root/reference/video_depth_anything unless noted).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

ENCODERS = {
    # dinov2.py:339-378, video_depth.py:53-56
    "vits": dict(D=384, depth=12, heads=6, taps=[2, 5, 8, 11]),
    "vitl": dict(D=1024, depth=24, heads=16, taps=[4, 11, 17, 23]),
}

# video_depth.py:29-33 ("infer settings, do not change")
INFER_LEN = 32
OVERLAP = 10
KEYFRAMES = [0, 12, 24, 25, 26, 27, 28, 29, 30, 31]
INTERP_LEN = 8


# --------------------------------------------------------------------------------------
# encoder
# --------------------------------------------------------------------------------------
def interpolate_pos_encoding(pos_embed: torch.Tensor, hp: int, wp: int) -> torch.Tensor:
    """dinov2.py:179-210.  pos_embed [1,1+37*37,D] -> [1,1+hp*wp,D].

    Identity only for the square 37x37 grid (:183-184); otherwise fp32 bicubic with
    scale_factor=((hp+0.1)/37,(wp+0.1)/37) (:194-205), i.e. PyTorch maps
    src = (dst+0.5)/scale_factor-0.5.
    """
    N = pos_embed.shape[1] - 1
    if hp * wp == N and hp == wp:
        return pos_embed
    pe = pos_embed.float()
    cls_pe, patch_pe = pe[:, 0], pe[:, 1:]
    D = pe.shape[-1]
    s = int(math.sqrt(N))
    sx, sy = float(hp + 0.1) / math.sqrt(N), float(wp + 0.1) / math.sqrt(N)
    patch_pe = F.interpolate(patch_pe.reshape(1, s, s, D).permute(0, 3, 1, 2),
                             scale_factor=(sx, sy), mode="bicubic", antialias=False)
    assert patch_pe.shape[-2] == hp and patch_pe.shape[-1] == wp
    patch_pe = patch_pe.permute(0, 2, 3, 1).reshape(1, -1, D)
    return torch.cat((cls_pe.unsqueeze(0), patch_pe), dim=1)


def prepare_tokens(sd, x: torch.Tensor) -> torch.Tensor:
    """dinov2.py:212-231 + dinov2_layers/patch_embed.py:69-82.  x [BT,3,H,W] -> [BT,1+hp*wp,D]."""
    BT, _, H, W = x.shape
    assert H % 14 == 0 and W % 14 == 0  # patch_embed.py:73-74
    t = F.conv2d(x, sd["pretrained.patch_embed.proj.weight"], sd["pretrained.patch_embed.proj.bias"], stride=14)
    t = t.flatten(2).transpose(1, 2)
    t = torch.cat((sd["pretrained.cls_token"].expand(BT, -1, -1), t), dim=1)
    return t + interpolate_pos_encoding(sd["pretrained.pos_embed"], H // 14, W // 14)


def vit_block(sd, prefix: str, x: torch.Tensor, heads: int) -> torch.Tensor:
    """dinov2_layers/block.py:82-107 (eval branch) with Attention.forward
    (dinov2_layers/attention.py:49-62), LayerScale (layer_scale.py:27-28), Mlp (mlp.py:35-41)."""
    B, N, C = x.shape
    h = F.layer_norm(x, (C,), sd[prefix + "norm1.weight"], sd[prefix + "norm1.bias"], 1e-6)
    qkv = F.linear(h, sd[prefix + "attn.qkv.weight"], sd[prefix + "attn.qkv.bias"])
    qkv = qkv.reshape(B, N, 3, heads, C // heads).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * (C // heads) ** -0.5, qkv[1], qkv[2]
    a = (q @ k.transpose(-2, -1)).softmax(dim=-1)
    h = (a @ v).transpose(1, 2).reshape(B, N, C)
    h = F.linear(h, sd[prefix + "attn.proj.weight"], sd[prefix + "attn.proj.bias"])
    x = x + h * sd[prefix + "ls1.gamma"]
    h = F.layer_norm(x, (C,), sd[prefix + "norm2.weight"], sd[prefix + "norm2.bias"], 1e-6)
    h = F.linear(h, sd[prefix + "mlp.fc1.weight"], sd[prefix + "mlp.fc1.bias"])
    h = F.gelu(h)
    h = F.linear(h, sd[prefix + "mlp.fc2.weight"], sd[prefix + "mlp.fc2.bias"])
    return x + h * sd[prefix + "ls2.gamma"]


def encoder_taps(sd, x: torch.Tensor, encoder: str, stages=None):
    """dinov2.py:297-321 (`get_intermediate_layers`, norm=True, return_class_token=True) via
    `_get_intermediate_layers_not_chunked` :271-281.  Returns 4 patch-token tensors [BT,hp*wp,D]."""
    cfg = ENCODERS[encoder]
    t = prepare_tokens(sd, x)
    if stages is not None:
        stages["tokens0"] = t
    taps = []
    for i in range(cfg["depth"]):
        t = vit_block(sd, f"pretrained.blocks.{i}.", t, cfg["heads"])
        if stages is not None:
            stages[f"block{i}"] = t
        if i in cfg["taps"]:
            n = F.layer_norm(t, (cfg["D"],), sd["pretrained.norm.weight"], sd["pretrained.norm.bias"], 1e-6)
            taps.append(n[:, 1:])
    return taps


# --------------------------------------------------------------------------------------
# motion module
# --------------------------------------------------------------------------------------
def temporal_attention(sd, p: str, h: torch.Tensor, T: int, heads: int = 8) -> torch.Tensor:
    """motion_module/motion_module.py:230-297 (ape) + motion_module/attention.py:182-211.
    h: [(b f), d, C] already layer-normed."""
    BF, d, C = h.shape
    b = BF // T
    h = h.reshape(b, T, d, C).permute(0, 2, 1, 3).reshape(b * d, T, C)      # "(b f) d c -> (b d) f c"
    h = h + sd[p + "pos_encoder.pe"][:, :T]                                   # :234-235, :197
    q = F.linear(h, sd[p + "to_q.weight"])
    k = F.linear(h, sd[p + "to_k.weight"])
    v = F.linear(h, sd[p + "to_v.weight"])
    dh = C // heads

    def split(t):  # reshape_heads_to_batch_dim, attention.py:93-98
        return t.reshape(b * d, T, heads, dh).permute(0, 2, 1, 3).reshape(b * d * heads, T, dh)

    q, k, v = split(q), split(k), split(v)
    s = torch.baddbmm(torch.zeros(1, dtype=q.dtype, device=q.device), q, k.transpose(-1, -2),
                      beta=0, alpha=dh ** -0.5)
    a = s.softmax(dim=-1)
    o = torch.bmm(a, v)
    o = o.reshape(b * d, heads, T, dh).permute(0, 2, 1, 3).reshape(b * d, T, C)
    o = F.linear(o, sd[p + "to_out.0.weight"], sd[p + "to_out.0.bias"])
    return o.reshape(b, d, T, C).permute(0, 2, 1, 3).reshape(BF, d, C)        # "(b d) f c -> (b f) d c"


def temporal_module(sd, m: int, x: torch.Tensor, T: int, stages=None) -> torch.Tensor:
    """motion_module.py:60-65 -> TemporalTransformer3DModel.forward :102-126 ->
    TemporalTransformerBlock.forward :164-177.  x: [(b f), C, h, w] (frame-major)."""
    p = f"head.motion_modules.{m}.temporal_transformer."
    BF, C, hh, ww = x.shape
    res = x
    h = F.group_norm(x, 32, sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-6)
    h = h.permute(0, 2, 3, 1).reshape(BF, hh * ww, C)
    h = F.linear(h, sd[p + "proj_in.weight"], sd[p + "proj_in.bias"])
    blk = p + "transformer_blocks.0."
    for a in (0, 1):
        n = F.layer_norm(h, (C,), sd[f"{blk}norms.{a}.weight"], sd[f"{blk}norms.{a}.bias"], 1e-5)
        h = temporal_attention(sd, f"{blk}attention_blocks.{a}.", n, T) + h
    n = F.layer_norm(h, (C,), sd[blk + "ff_norm.weight"], sd[blk + "ff_norm.bias"], 1e-5)
    g = F.linear(n, sd[blk + "ff.net.0.proj.weight"], sd[blk + "ff.net.0.proj.bias"])
    a_, gate = g.chunk(2, dim=-1)                                              # attention.py:382-384
    n = a_ * F.gelu(gate)
    h = F.linear(n, sd[blk + "ff.net.2.weight"], sd[blk + "ff.net.2.bias"]) + h
    h = F.linear(h, sd[p + "proj_out.weight"], sd[p + "proj_out.bias"])
    h = h.reshape(BF, hh, ww, C).permute(0, 3, 1, 2)
    return h + res


# --------------------------------------------------------------------------------------
# DPT head
# --------------------------------------------------------------------------------------
def rcu(sd, p: str, x: torch.Tensor) -> torch.Tensor:
    """util/blocks.py:68-91 (non-inplace ReLU; skip adds the pre-activation x)."""
    o = F.conv2d(F.relu(x), sd[p + "conv1.weight"], sd[p + "conv1.bias"], padding=1)
    o = F.conv2d(F.relu(o), sd[p + "conv2.weight"], sd[p + "conv2.bias"], padding=1)
    return o + x


def fusion(sd, r: int, x0: torch.Tensor, x1, size=None) -> torch.Tensor:
    """util/blocks.py:135-162 (FeatureFusionBlock.forward)."""
    p = f"head.scratch.refinenet{r}."
    out = x0
    if x1 is not None:
        out = out + rcu(sd, p + "resConfUnit1.", x1)
    out = rcu(sd, p + "resConfUnit2.", out)
    if size is None:
        out = F.interpolate(out, scale_factor=2, mode="bilinear", align_corners=True)
    else:
        out = F.interpolate(out, size=size, mode="bilinear", align_corners=True)
    return F.conv2d(out, sd[p + "out_conv.weight"], sd[p + "out_conv.bias"])


def dpt_head(sd, taps, hp: int, wp: int, T: int, stages=None) -> torch.Tensor:
    """dpt_temporal.py:53-114 (weights built by dpt.py:47-124).  The 4-frame micro-batching
    (:88-114) is a memory knob only and is not restated."""
    h = "head."
    out = []
    for i, x in enumerate(taps):
        BT, _, D = x.shape
        x = x.permute(0, 2, 1).reshape(BT, D, hp, wp)
        x = F.conv2d(x, sd[f"{h}projects.{i}.weight"], sd[f"{h}projects.{i}.bias"])
        if i == 0:
            x = F.conv_transpose2d(x, sd[h + "resize_layers.0.weight"], sd[h + "resize_layers.0.bias"], stride=4)
        elif i == 1:
            x = F.conv_transpose2d(x, sd[h + "resize_layers.1.weight"], sd[h + "resize_layers.1.bias"], stride=2)
        elif i == 3:
            x = F.conv2d(x, sd[h + "resize_layers.3.weight"], sd[h + "resize_layers.3.bias"], stride=2, padding=1)
        out.append(x)
    l1, l2, l3, l4 = out
    if stages is not None:
        stages.update(layer_1=l1, layer_2=l2, layer_3=l3, layer_4=l4)
    l3 = temporal_module(sd, 0, l3, T)
    l4 = temporal_module(sd, 1, l4, T)
    if stages is not None:
        stages.update(mm0=l3, mm1=l4)
    l1r = F.conv2d(l1, sd[h + "scratch.layer1_rn.weight"], padding=1)
    l2r = F.conv2d(l2, sd[h + "scratch.layer2_rn.weight"], padding=1)
    l3r = F.conv2d(l3, sd[h + "scratch.layer3_rn.weight"], padding=1)
    l4r = F.conv2d(l4, sd[h + "scratch.layer4_rn.weight"], padding=1)
    if stages is not None:
        stages.update(layer_1_rn=l1r, layer_2_rn=l2r, layer_3_rn=l3r, layer_4_rn=l4r)
    p4 = fusion(sd, 4, l4r, None, size=l3r.shape[2:])
    if stages is not None:
        stages["path_4_pre"] = p4
    p4 = temporal_module(sd, 2, p4, T)
    p3 = fusion(sd, 3, p4, l3r, size=l2r.shape[2:])
    if stages is not None:
        stages.update(path_4=p4, path_3_pre=p3)
    p3 = temporal_module(sd, 3, p3, T)
    p2 = fusion(sd, 2, p3, l2r, size=l1r.shape[2:])
    p1 = fusion(sd, 1, p2, l1r, None)
    if stages is not None:
        stages.update(path_3=p3, path_2=p2, path_1=p1)
    o = F.conv2d(p1, sd[h + "scratch.output_conv1.weight"], sd[h + "scratch.output_conv1.bias"], padding=1)
    if stages is not None:
        stages["output_conv1"] = o
    o = F.interpolate(o, (hp * 14, wp * 14), mode="bilinear", align_corners=True)
    o = F.relu(F.conv2d(o, sd[h + "scratch.output_conv2.0.weight"], sd[h + "scratch.output_conv2.0.bias"], padding=1))
    o = F.relu(F.conv2d(o, sd[h + "scratch.output_conv2.2.weight"], sd[h + "scratch.output_conv2.2.bias"]))
    return o


@torch.no_grad()
def forward(sd, x: torch.Tensor, encoder: str, stages=None) -> torch.Tensor:
    """video_depth.py:89-164 (oracle twin metric_depth/video_depth_anything/video_depth.py:58-65).
    x [B,T,3,H,W] -> depth [B,T,H,W].  B must be 1 per temporal group like the reference."""
    B, T, C, H, W = x.shape
    hp, wp = H // 14, W // 14
    taps = encoder_taps(sd, x.flatten(0, 1), encoder, stages)
    if stages is not None:
        for i, t in enumerate(taps):
            stages[f"tap{i}"] = t
    d = dpt_head(sd, taps, hp, wp, T, stages)
    d = F.interpolate(d, size=(H, W), mode="bilinear", align_corners=True)
    d = F.relu(d)
    return d.squeeze(1).unflatten(0, (B, T))


# --------------------------------------------------------------------------------------
# long-video driver (host numpy arithmetic, restated with numpy like the reference)
# --------------------------------------------------------------------------------------
def window_source_indices(n_frames: int):
    """Closed form of the inputs of every window of video_depth.py:187-201:
    window 0 = frames 0..31; window k>=1 = [0, 22k-10, 22k+2 .. 22k+31], clipped to n-1."""
    step = INFER_LEN - OVERLAP
    wins = []
    k = 0
    for frame_id in range(0, n_frames, step):
        if k == 0:
            idx = list(range(INFER_LEN))
        else:
            idx = [0, step * k - 10] + [step * k + 2 + j for j in range(30)]
        wins.append([min(i, n_frames - 1) for i in idx])
        k += 1
    return wins


def window_source_indices_literal(n_frames: int):
    """The same table produced by literally replaying video_depth.py:187-201 on index-valued
    frames (pad, slice 32, overwrite the first 10 with the previous window's KEYFRAMES)."""
    step = INFER_LEN - OVERLAP
    frame_list = list(range(n_frames))
    append = (step - (n_frames % step)) % step + (INFER_LEN - step)
    frame_list = frame_list + [frame_list[-1]] * append
    wins, pre = [], None
    for frame_id in range(0, n_frames, step):
        cur = [frame_list[frame_id + i] for i in range(INFER_LEN)]
        if pre is not None:
            for j, kf in enumerate(KEYFRAMES):
                cur[j] = pre[kf]
        wins.append(cur)
        pre = cur
    return wins


def compute_scale_and_shift(prediction, target):
    """utils/util.py:40-62 with the all-ones mask the driver passes (video_depth.py:230-232)."""
    prediction = prediction.astype(np.float32)
    target = target.astype(np.float32)
    mask = np.ones_like(prediction, dtype=np.float32)
    a_00 = np.sum(mask * prediction * prediction)
    a_01 = np.sum(mask * prediction)
    a_11 = np.sum(mask)
    b_0 = np.sum(mask * prediction * target)
    b_1 = np.sum(mask * target)
    x_0, x_1 = 1, 0
    det = a_00 * a_11 - a_01 * a_01
    if det != 0:
        x_0 = (a_11 * b_0 - a_01 * b_1) / det
        x_1 = (-a_01 * b_0 + a_00 * b_1) / det
    return x_0, x_1


def align_windows(depth_list, n_frames: int, mode: str = "affine"):
    """video_depth.py:216-254 (+ utils/util.py:65-74).  depth_list: list of K*32 [H0,W0] float32
    arrays (raw per-window depths, already resized).  mode 'identity' is the metric variant
    (metric_depth/video_depth_anything/video_depth.py:132: scale, shift = 1., 0.)."""
    aligned, ref_align = [], []
    align_len = OVERLAP - INTERP_LEN
    kf_align_list = KEYFRAMES[:align_len]
    for frame_id in range(0, len(depth_list), INFER_LEN):
        if len(aligned) == 0:
            aligned += depth_list[:INFER_LEN]
            for kf in kf_align_list:
                ref_align.append(depth_list[frame_id + kf])
        else:
            curr = [depth_list[frame_id + i] for i in range(len(kf_align_list))]
            if mode == "affine":
                scale, shift = compute_scale_and_shift(np.concatenate(curr), np.concatenate(ref_align))
            else:
                scale, shift = 1.0, 0.0
            pre = aligned[-INTERP_LEN:]
            post = [depth_list[frame_id + align_len + i] * scale + shift for i in range(INTERP_LEN)]
            for pd in post:
                pd[pd < 0] = 0
            step = 1.0 / (INTERP_LEN - 1)
            w = [0.0] + [i * step for i in range(1, INTERP_LEN - 1)] + [1.0]
            aligned[-INTERP_LEN:] = [pre[i] * (1 - w[i]) + post[i] * w[i] for i in range(INTERP_LEN)]
            for i in range(OVERLAP, INFER_LEN):
                nd = depth_list[frame_id + i] * scale + shift
                nd[nd < 0] = 0
                aligned.append(nd)
            ref_align = ref_align[:1]
            for kf in kf_align_list[1:]:
                nd = depth_list[frame_id + kf] * scale + shift
                nd[nd < 0] = 0
                ref_align.append(nd)
    return np.stack(aligned[:n_frames], axis=0)


def get_resize_hw(h0: int, w0: int, input_size: int):
    """util/transform.py:62-107 with the driver's arguments (video_depth.py:173-185):
    keep_aspect_ratio, lower_bound, multiple of 14; plus the aspect guard video_depth.py:167-171."""
    ratio = max(h0, w0) / min(h0, w0)
    if ratio > 1.78:
        input_size = int(input_size * 1.777 / ratio)
        input_size = round(input_size / 14) * 14
    sh, sw = input_size / h0, input_size / w0
    if sw > sh:
        sh = sw
    else:
        sw = sh

    def constrain(x, min_val):
        y = (np.round(x / 14) * 14).astype(int)
        if y < min_val:
            y = (np.ceil(x / 14) * 14).astype(int)
        return int(y)

    return constrain(sh * h0, input_size), constrain(sw * w0, input_size)


def preprocess_frame(frame_u8: np.ndarray, input_size: int) -> np.ndarray:
    """util/transform.py:109-158 as composed in video_depth.py:173-185,198: /255 -> cv2 INTER_CUBIC
    resize -> (x-mean)/std in float64 -> CHW float32."""
    import cv2
    h0, w0 = frame_u8.shape[:2]
    nh, nw = get_resize_hw(h0, w0, input_size)
    img = frame_u8.astype(np.float32) / 255.0
    img = cv2.resize(img, (nw, nh), interpolation=cv2.INTER_CUBIC)
    img = (img - [0.485, 0.456, 0.406]) / [0.229, 0.224, 0.225]
    return np.ascontiguousarray(np.transpose(img, (2, 0, 1))).astype(np.float32)


def _cubic_coeffs_f32(x: np.ndarray) -> np.ndarray:
    """OpenCV's interpolateCubic (A = -0.75) evaluated in float32, as cv2.resize does for CV_32F images."""
    f = np.float32
    A, x = f(-0.75), x.astype(np.float32)
    c0 = ((A * (x + f(1)) - f(5) * A) * (x + f(1)) + f(8) * A) * (x + f(1)) - f(4) * A
    c1 = ((A + f(2)) * x - (A + f(3))) * x * x + f(1)
    c2 = ((A + f(2)) * (f(1) - x) - (A + f(3))) * (f(1) - x) * (f(1) - x) + f(1)
    c3 = f(1) - c0 - c1 - c2
    return np.stack([c0, c1, c2, c3], -1).astype(np.float32)


def resize_cubic_opencv(img: np.ndarray, nw: int, nh: int) -> np.ndarray:
    """cv2.resize(img float32 [H,W,C], (nw, nh), INTER_CUBIC) restated from OpenCV's published generic algorithm
    (third-party dependency of util/transform.py:109-111; opencv-python 4.13 in this image): source coordinate
    (d + 0.5) * (1 / (dst/src)) - 0.5 rounded to float32, taps floor-1..floor+2 with replicated borders, horizontal
    pass then vertical pass, float32 throughout; same-size = identity.  Pinned in tests/test_oracle_golden.py
    against cv2 with IPP switched off (<= 1 float32 ulp of 1.0); cv2's default IPP build of the same resize differs
    from this generic path by up to ~1e-4 (implementation-defined, the tolerance the device kernel is held to
    against stock cv2)."""
    H, W = img.shape[:2]
    if (H, W) == (nh, nw):
        return img.astype(np.float32).copy()

    def axis(n, src):
        scale = 1.0 / (n / src)
        f = ((np.arange(n) + 0.5) * scale - 0.5).astype(np.float32)
        s = np.floor(f).astype(np.int64)
        c = _cubic_coeffs_f32(f - s.astype(np.float32))
        return np.clip(s[:, None] + np.arange(-1, 3)[None], 0, src - 1), c

    ix, cx = axis(nw, W)
    iy, cy = axis(nh, H)
    img = img.astype(np.float32)
    rows = np.zeros((H, nw, img.shape[2]), np.float32)
    for k in range(4):
        rows = rows + img[:, ix[:, k], :] * cx[None, :, k, None]
    out = np.zeros((nh, nw, img.shape[2]), np.float32)
    for k in range(4):
        out = out + rows[iy[:, k]] * cy[:, k, None, None]
    return out


def preprocess_frame_generic(frame_u8: np.ndarray, input_size: int) -> np.ndarray:
    """preprocess_frame with the resize done by resize_cubic_opencv (no cv2 call)."""
    h0, w0 = frame_u8.shape[:2]
    nh, nw = get_resize_hw(h0, w0, input_size)
    img = resize_cubic_opencv(frame_u8.astype(np.float32) / 255.0, nw, nh)
    img = (img - [0.485, 0.456, 0.406]) / [0.229, 0.224, 0.225]
    return np.ascontiguousarray(np.transpose(img, (2, 0, 1))).astype(np.float32)


@torch.no_grad()
def infer_video_depth(sd, frames: np.ndarray, encoder: str, input_size: int = 518,
                      mode: str = "affine", forward_fn=None):
    """video_depth.py:166-254.  frames uint8 [N,H0,W0,3] -> float32 [N,H0,W0]."""
    n = frames.shape[0]
    h0, w0 = frames.shape[1:3]
    if forward_fn is None:
        forward_fn = lambda x: forward(sd, x, encoder)
    pre = {}
    depth_list = []
    for idx in window_source_indices(n):
        for i in set(idx):
            if i not in pre:
                pre[i] = torch.from_numpy(preprocess_frame(frames[i], input_size))
        x = torch.stack([pre[i] for i in idx]).unsqueeze(0)
        d = forward_fn(x).float()
        d = F.interpolate(d.flatten(0, 1).unsqueeze(1), size=(h0, w0), mode="bilinear", align_corners=True)
        depth_list += [d[i, 0].cpu().numpy() for i in range(d.shape[0])]
    return align_windows(depth_list, n, mode)


def rel_err(a: torch.Tensor, ref: torch.Tensor):
    """Tolerance metric of SURVEY.md §8(d): |a-ref| / max(|ref|, 1e-3*max|ref|); returns (max, p99.9, mean)."""
    a, ref = a.double().flatten(), ref.double().flatten()
    den = ref.abs().clamp_min(1e-3 * ref.abs().max())
    r = (a - ref).abs() / den
    k = max(1, int(r.numel() * 0.999))
    p999 = r.kthvalue(k).values.item() if r.numel() < (1 << 24) else float(np.quantile(r.cpu().numpy(), 0.999))
    return r.max().item(), p999, r.mean().item()
