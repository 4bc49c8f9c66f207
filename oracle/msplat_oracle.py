"""CPU restatement of the msplat render path -- TEST INFRASTRUCTURE ONLY.

This module is the *oracle* for ``pointrix_b200``: a plain-PyTorch (device
agnostic, fp32, autograd-differentiable) restatement of the six msplat
operators and of the ``MsplatRender.render_iter`` glue.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs may import it; the product package never does.

Every function cites the reference lines it follows (paths relative to
``/root/reference``).  Backward passes are obtained with autograd through the
forward restatement: each reference backward kernel is the analytic derivative
of its forward (the reference's own tests check exactly that,
``msplat/test/test_*.py``), with the discrete decisions (culling, alpha
thresholds, early termination) held fixed.

Parity status
-------------
* integer work that is pure integer / IEEE arithmetic (tile rectangles, 64-bit
  keys, stable key sort, tile ranges) is bit-exact here and is pinned by the
  reference's known-answer test (``msplat/test/test_sort_gaussian.py:8-52``)
  in ``tests/test_oracle.py``.
* ``radius``/``tiles`` depend on the GPU's ``MUFU.RCP/SQRT`` approximations
  (``--use_fast_math`` build); on CPU they can only be reproduced up to rare
  +-1 flips at ``ceil`` boundaries.  The bit-exact checker for those is the
  compiled reference itself (``oracle/_ref``, see ``oracle/build_ref.py``) and
  the golden vectors it generated on a B200 (``tests/golden``).
* floating outputs: tolerance oracles (image max-abs 1e-4, gradients rel 1e-3).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch

BLOCK = 16  # msplat/msplat/include/config.h:7-8


# ----------------------------------------------------------------------------
# a6  project_point            msplat/msplat/src/project_point.cu:27-56
# ----------------------------------------------------------------------------
def camera_transform(xyz: torch.Tensor, extr: torch.Tensor) -> torch.Tensor:
    """t = E[3x4] . [p;1]  (project_point.cu:28-30, ewa_project.cu:35-39)."""
    e = extr.reshape(3, 4)
    return xyz @ e[:, :3].t() + e[:, 3]


def project_point(xyz, intr, extr, W: int, H: int, nearest: float = 0.0, extent: float = 1.3):
    """uv = f * t.xy / (t.z + 1e-7) + c - 0.5 ; depth = t.z ; culled => zeros.

    project_point.cu:27-56.  The reciprocal is evaluated in fp64 and rounded to
    fp32 (``float norm1 = 1.0 / (tmp.z + 1e-7)``, :31).
    """
    intr = intr.reshape(-1)
    t = camera_transform(xyz, extr)
    tz = t[:, 2]
    norm1 = (1.0 / (tz.double() + 1e-7)).to(xyz.dtype)
    u = intr[0] * t[:, 0] * norm1 + intr[2] - 0.5
    v = intr[1] * t[:, 1] * norm1 + intr[3] - 0.5
    cull = torch.zeros_like(tz, dtype=torch.bool)
    if nearest > 0:
        cull = cull | (tz <= nearest)
    if extent > 0:
        f32 = torch.float32
        x_min = torch.tensor((1 - extent), dtype=f32) * W * 0.5
        x_max = torch.tensor((1 + extent), dtype=f32) * W * 0.5
        y_min = torch.tensor((1 - extent), dtype=f32) * H * 0.5
        y_max = torch.tensor((1 + extent), dtype=f32) * H * 0.5
        cull = cull | (u < x_min) | (u > x_max) | (v < y_min) | (v > y_max)
    # NaN coordinates compare false everywhere => not culled in the reference either
    keep = ~cull
    uv = torch.stack([u, v], dim=-1)
    uv = torch.where(keep[:, None], uv, torch.zeros_like(uv))
    depth = torch.where(keep, tz, torch.zeros_like(tz))[:, None]
    return uv, depth


# ----------------------------------------------------------------------------
# a7  compute_cov3d            msplat/msplat/src/compute_cov3d.cu:24-58
# ----------------------------------------------------------------------------
def quat_to_rotmat_glm(q: torch.Tensor) -> torch.Tensor:
    """GLM column-major matrix Rg[c][k] of compute_cov3d.cu:24-40 returned as a
    [P,3(c),3(k)] tensor (columns first)."""
    r, x, y, z = q.unbind(-1)
    c0 = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y)], -1)
    c1 = torch.stack([2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x)], -1)
    c2 = torch.stack([2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], -1)
    return torch.stack([c0, c1, c2], dim=-2)


def compute_cov3d(scales, uquats, visible: Optional[torch.Tensor] = None):
    """Sigma = M^T M with GLM ``M = S * R`` => M[c][k] = s_k * Rg[c][k]; upper
    triangle [S00,S01,S02,S11,S12,S22] (compute_cov3d.cu:42-58)."""
    Rg = quat_to_rotmat_glm(uquats)  # [P, c, k]
    M = Rg * scales[:, None, :]  # M[c][k] = s_k Rg[c][k]
    Sigma = M @ M.transpose(-1, -2)  # Sigma[c][r] = sum_k M[r][k] M[c][k]
    cov = torch.stack(
        [Sigma[:, 0, 0], Sigma[:, 0, 1], Sigma[:, 0, 2], Sigma[:, 1, 1], Sigma[:, 1, 2], Sigma[:, 2, 2]], -1
    )
    if visible is not None:
        cov = torch.where(visible.reshape(-1, 1).bool(), cov, torch.zeros_like(cov))
    return cov


# ----------------------------------------------------------------------------
# a9  get_rect                 msplat/msplat/include/utils.h:17-37
# ----------------------------------------------------------------------------
def get_rect(uv: torch.Tensor, radius: torch.Tensor, W: int, H: int):
    """Tile rectangle [min,max) per Gaussian; float ops in the reference's order,
    truncation toward zero, clamp to the grid.  Pure IEEE fp32 => bit-exact."""
    gx, gy = (W + BLOCK - 1) // BLOCK, (H + BLOCK - 1) // BLOCK
    rf = radius.to(torch.float32)
    u, v = uv[:, 0].float(), uv[:, 1].float()
    s = torch.tensor(1.0 / BLOCK, dtype=torch.float32)

    def lo(c, g):
        return ((c - rf) * s).nan_to_num(0.0, 2.0e9, -2.0e9).clamp(-2.0e9, 2.0e9).to(torch.int64).clamp(0, g)

    def hi(c, g):
        return ((((c + rf) + float(BLOCK)) - 1.0) * s).nan_to_num(0.0, 2.0e9, -2.0e9).clamp(-2.0e9, 2.0e9).to(torch.int64).clamp(0, g)

    return lo(u, gx), lo(v, gy), hi(u, gx), hi(v, gy)


# ----------------------------------------------------------------------------
# a8  ewa_project              msplat/msplat/src/ewa_project.cu:31-82
# ----------------------------------------------------------------------------
def ewa_cov2d(xyz, cov3d, intr, extr):
    """cov2D = T Sigma T^T with T = J W, J without frustum clamp (ewa_project.cu:41-56).
    Returns (a, b, c) = (cov00 + 0.3, cov01, cov11 + 0.3) and t."""
    intr = intr.reshape(-1)
    e = extr.reshape(3, 4)
    fx, fy = intr[0], intr[1]
    t = camera_transform(xyz, extr)
    tx, ty, tz = t.unbind(-1)
    rz = 1.0 / tz
    rz2 = 1.0 / (tz * tz)
    zero = torch.zeros_like(tz)
    # rows of the 2x3 Jacobian
    J0 = torch.stack([fx * rz, zero, -(fx * tx) * rz2], -1)
    J1 = torch.stack([zero, fy * rz, -(fy * ty) * rz2], -1)
    J = torch.stack([J0, J1], -2)  # [P,2,3]
    Tm = J @ e[:, :3]  # [P,2,3]
    V = torch.stack(
        [
            torch.stack([cov3d[:, 0], cov3d[:, 1], cov3d[:, 2]], -1),
            torch.stack([cov3d[:, 1], cov3d[:, 3], cov3d[:, 4]], -1),
            torch.stack([cov3d[:, 2], cov3d[:, 4], cov3d[:, 5]], -1),
        ],
        -2,
    )
    cov2d = Tm @ V @ Tm.transpose(-1, -2)
    a = cov2d[:, 0, 0] + 0.3
    b = cov2d[:, 0, 1]
    c = cov2d[:, 1, 1] + 0.3
    return a, b, c, t


def ewa_project(xyz, cov3d, intr, extr, uv, W: int, H: int, visible: Optional[torch.Tensor] = None):
    """conic, radius(int32), tiles(int32)  (ewa_project.cu:58-82)."""
    P = xyz.shape[0]
    if visible is None:
        visible = torch.ones(P, dtype=torch.bool, device=xyz.device)
    visible = visible.reshape(-1).bool()
    a, b, c, _ = ewa_cov2d(xyz, cov3d, intr, extr)
    det = a * c - b * b
    mid = 0.5 * (a + c)
    s = torch.sqrt(torch.clamp(mid * mid - det, min=0.1))
    lam = torch.maximum(mid + s, mid - s)
    rad_f = torch.ceil(3.0 * torch.sqrt(lam))
    ok = visible & (det != 0) & torch.isfinite(rad_f)
    radius = torch.where(ok, rad_f, torch.zeros_like(rad_f)).clamp(-2.0e9, 2.0e9).to(torch.int32)
    x0, y0, x1, y1 = get_rect(uv.detach(), radius, W, H)
    tiles = ((y1 - y0) * (x1 - x0)).to(torch.int32)
    ok = ok & (tiles != 0)
    det_safe = torch.where(ok, det, torch.ones_like(det))
    conic = torch.stack([c / det_safe, -b / det_safe, a / det_safe], -1)
    conic = torch.where(ok[:, None], conic, torch.zeros_like(conic))
    radius = torch.where(ok, radius, torch.zeros_like(radius))
    tiles = torch.where(ok, tiles, torch.zeros_like(tiles))
    return conic, radius, tiles


# ----------------------------------------------------------------------------
# a10 sort_gaussian            msplat/msplat/sort_gaussian.py:42-52,
#                              msplat/msplat/src/sort_gaussian.cu:29-42,53-70
# ----------------------------------------------------------------------------
def gaussian_keys(uv, depth, W: int, H: int, radius, tiles):
    """(keys int64 [N], gaussian_idx int32 [N]) in emission order: Gaussians by
    ascending id, each one's tile rect row-major; key = tile<<32 | sign-extended
    float bits of depth (sort_gaussian.cu:29-42)."""
    gx = (W + BLOCK - 1) // BLOCK
    radius = radius.reshape(-1)
    x0, y0, x1, y1 = get_rect(uv, radius, W, H)
    live = radius > 0
    w = torch.where(live, x1 - x0, torch.zeros_like(x0))
    h = torch.where(live, y1 - y0, torch.zeros_like(y0))
    cnt = w * h
    # the reference places Gaussian g at cumsum(tiles)[g-1]; tiles == cnt for
    # every live Gaussian (ewa_project.cu:81 computes the same rect)
    offs_in = torch.cumsum(tiles.reshape(-1).to(torch.int64), 0)
    N = int(offs_in[-1]) if offs_in.numel() else 0
    start = offs_in - tiles.reshape(-1).to(torch.int64)
    gid = torch.repeat_interleave(torch.arange(radius.numel(), device=uv.device), cnt)
    local = torch.arange(gid.numel(), device=uv.device) - torch.repeat_interleave(cnt.cumsum(0) - cnt, cnt)
    ww = w[gid]
    ty = y0[gid] + local // torch.clamp(ww, min=1)
    tx = x0[gid] + local % torch.clamp(ww, min=1)
    tile_id = ty * gx + tx
    dbits = depth.reshape(-1).contiguous().view(torch.int32).to(torch.int64)[gid]
    keys_e = (tile_id << 32) | dbits  # sign-extended OR, as the reference's (int64)(int)
    keys = torch.zeros(N, dtype=torch.int64, device=uv.device)
    idx = torch.zeros(N, dtype=torch.int32, device=uv.device)
    pos = start[gid] + local
    keys[pos] = keys_e
    idx[pos] = gid.to(torch.int32)
    return keys, idx


def tile_ranges(keys_sorted: torch.Tensor, num_tiles: int) -> torch.Tensor:
    """[first,last+1) per tile from key>>32 boundaries; untouched tiles (0,0)
    (sort_gaussian.cu:53-70)."""
    rng = torch.zeros(num_tiles, 2, dtype=torch.int32, device=keys_sorted.device)
    N = keys_sorted.numel()
    if N == 0:
        return rng
    t = (keys_sorted >> 32).to(torch.int64)
    first = torch.ones(N, dtype=torch.bool, device=t.device)
    first[1:] = t[1:] != t[:-1]
    starts = torch.nonzero(first).reshape(-1)
    ends = torch.cat([starts[1:], torch.tensor([N], device=t.device)])
    rng[t[starts], 0] = starts.to(torch.int32)
    rng[t[starts], 1] = ends.to(torch.int32)
    return rng


def sort_gaussian(uv, depth, W: int, H: int, radius, tiles):
    """idx_sorted int32 [N], tile_range int32 [tiles,2]; ascending int64 key sort,
    ties in emission order (stable) (sort_gaussian.py:42-52)."""
    keys, idx = gaussian_keys(uv, depth, W, H, radius, tiles)
    ks, perm = torch.sort(keys, stable=True)
    idx_sorted = idx[perm]
    gx, gy = (W + BLOCK - 1) // BLOCK, (H + BLOCK - 1) // BLOCK
    return idx_sorted, tile_ranges(ks, gx * gy)


# ----------------------------------------------------------------------------
# a5  compute_sh               msplat/msplat/src/compute_sh.cu:17-35,116-164 (deg<=3)
#                              and :166-503 (deg 4..10: c * Q_n^m(z) * {A_m,B_m}(x,y))
# ----------------------------------------------------------------------------
def _sh_norm(n: int, m: int) -> float:
    return math.sqrt((2 * n + 1) / (4 * math.pi) * math.factorial(n - m) / math.factorial(n + m))


def sh_bases(D: int, dirs: torch.Tensor) -> torch.Tensor:
    """Real SH basis [P, D], index n*n+n+m, Condon-Shortley phase kept
    ((-1)^m), the 3DGS sign convention.  Degrees <= 3 use the homogeneous
    polynomial forms of compute_sh.cu:116-164; degrees >= 4 are
    ``(-1)^m sqrt2 N_n^m Q_n^m(z) {A_m,B_m}(x,y)`` with Q_n^m = d^m P_n/dz^m and
    A_m + i B_m = (x + i y)^m, which is the factorisation the reference's
    polynomials (compute_sh.cu:166-503) expand to."""
    deg = int(round(math.sqrt(D))) - 1
    assert (deg + 1) ** 2 == D and 0 <= deg <= 10, "D must be a square <= 121"
    x, y, z = dirs.unbind(-1)
    out = [torch.full_like(x, 0.28209479177387814)]
    if deg >= 1:
        C1 = 0.4886025119029199
        out += [-C1 * y, C1 * z, -C1 * x]
    if deg >= 2:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        out += [
            1.0925484305920792 * xy,
            -1.0925484305920792 * yz,
            0.31539156525252005 * (2.0 * zz - xx - yy),
            -1.0925484305920792 * xz,
            0.5462742152960396 * (xx - yy),
        ]
    if deg >= 3:
        out += [
            -0.5900435899266435 * y * (3.0 * xx - yy),
            2.890611442640554 * xy * z,
            -0.4570457994644658 * y * (4.0 * zz - xx - yy),
            0.3731763325901154 * z * (2.0 * zz - 3.0 * xx - 3.0 * yy),
            -0.4570457994644658 * x * (4.0 * zz - xx - yy),
            1.445305721320277 * z * (xx - yy),
            -0.5900435899266435 * x * (xx - 3.0 * yy),
        ]
    if deg >= 4:
        # A_m, B_m
        A = [torch.ones_like(x)]
        B = [torch.zeros_like(x)]
        for m in range(1, deg + 1):
            A.append(x * A[m - 1] - y * B[m - 1])
            B.append(x * B[m - 1] + y * A[m - 1])
        # Q[n][m]
        Q = [[None] * (deg + 2) for _ in range(deg + 1)]
        dfact = 1.0
        for m in range(0, deg + 1):
            if m > 0:
                dfact *= 2 * m - 1
            Q[m][m] = torch.full_like(z, dfact)
            if m + 1 <= deg:
                Q[m + 1][m] = (2 * m + 1) * z * Q[m][m]
            for n in range(m + 2, deg + 1):
                Q[n][m] = ((2 * n - 1) * z * Q[n - 1][m] - (n + m - 1) * Q[n - 2][m]) / (n - m)
        for n in range(4, deg + 1):
            row = [None] * (2 * n + 1)
            row[n] = _sh_norm(n, 0) * Q[n][0]
            for m in range(1, n + 1):
                c = (-1) ** m * math.sqrt(2.0) * _sh_norm(n, m)
                row[n + m] = c * Q[n][m] * A[m]
                row[n - m] = c * Q[n][m] * B[m]
            out += row
    return torch.stack(out, dim=-1)


def compute_sh(shs, view_dirs, visible: Optional[torch.Tensor] = None):
    """value[p,c] = sum_d basis_d(dir_p) shs[p,c,d]; rows with !visible are 0
    (compute_sh.cu:1600-1637).  No +0.5 / clamp here (done by the caller)."""
    D = shs.shape[-1]
    bases = sh_bases(D, view_dirs)
    val = (bases[:, None, :] * shs).sum(-1)
    if visible is not None:
        val = torch.where(visible.reshape(-1, 1).bool(), val, torch.zeros_like(val))
    return val


# ----------------------------------------------------------------------------
# a11/a12 alpha_blending       msplat/msplat/src/alpha_blending.cu:54-109 (fwd), 155-244 (bwd)
# ----------------------------------------------------------------------------
def alpha_blending(
    uv, conic, opacity, feature, idx_sorted, tile_range, bg: float, W: int, H: int,
    ndc: Optional[torch.Tensor] = None, return_aux: bool = False, max_elems: int = 1 << 24,
    dL_dout: Optional[torch.Tensor] = None, tile_mask: Optional[torch.Tensor] = None,
):
    """Front-to-back blend per 16x16 tile, vectorised over (tiles-in-chunk, 256
    pixels, list length).  Skip rules (alpha_blending.cu:80-94): power > 0;
    alpha = min(0.99, op*exp(power)) < 1/255; termination when T(1-alpha) < 1e-4
    (that Gaussian is NOT blended).  out = F + T_final * bg for every channel.

    ``ndc`` (if given, requires_grad) receives dL/duv * (W/2, H/2) exactly as
    msplat/msplat/alpha_blending.py:106-110 does: it enters the graph through a
    zero-valued term whose Jacobian is that scale.

    ``dL_dout`` ([C,H,W], optional): bounded-memory backward.  Every tile chunk is back-propagated as
    soon as it is blended (``autograd.backward(chunk, dL_dout[chunk])``, accumulating into the ``.grad`` of
    whatever leaves ``uv / conic / opacity / feature`` hang off) and its graph is dropped, so the
    [tiles, 256, list] intermediates of only ONE chunk are alive at a time; the returned image is
    detached.  Used by :func:`render_step` for the full-size CPU baseline (1 M Gaussians at 1080p would
    otherwise keep > 100 GB of autograd state).

    ``tile_mask`` (bool [tiles], optional): blend only the marked tiles (the others stay background) --
    the bounded SAMPLE of a full-size view that bench.py's ``cpu_baseline`` leg times.
    """
    C = feature.shape[1]
    gx, gy = (W + BLOCK - 1) // BLOCK, (H + BLOCK - 1) // BLOCK
    dev = uv.device
    if ndc is not None:
        scale = torch.tensor([0.5 * W, 0.5 * H], dtype=uv.dtype, device=dev)
        # value-neutral link: uv_eff == uv, d uv_eff / d ndc = scale
        uv = uv + (ndc - ndc.detach()) * scale
    tr = tile_range.to(torch.int64)
    counts = tr[:, 1] - tr[:, 0]
    if tile_mask is not None:
        counts = torch.where(tile_mask.to(dev), counts, torch.zeros_like(counts))
    out = torch.zeros(gy * BLOCK, gx * BLOCK, C, dtype=feature.dtype, device=dev)
    final_T = torch.ones(gy * BLOCK, gx * BLOCK, dtype=feature.dtype, device=dev)
    ncontrib = torch.zeros(gy * BLOCK, gx * BLOCK, dtype=torch.int32, device=dev)
    out = out + bg  # empty tiles: F = 0, T = 1
    px = torch.arange(BLOCK, device=dev, dtype=torch.float32)
    lx = px.repeat(BLOCK)  # pixel x within tile, row-major
    ly = px.repeat_interleave(BLOCK)
    order = torch.argsort(counts, stable=True)
    order = order[counts[order] > 0]
    pieces_out, pieces_T, pieces_n, pieces_tile = [], [], [], []
    i = 0
    idx_sorted = idx_sorted.to(torch.int64)
    dL_pad = None
    if dL_dout is not None:  # [C,H,W] -> tile-padded [gy*16, gx*16, C]
        dL_pad = torch.zeros(gy * BLOCK, gx * BLOCK, C, dtype=feature.dtype, device=dev)
        dL_pad[:H, :W] = dL_dout.detach().permute(1, 2, 0)
    while i < order.numel():
        # chunk of tiles with similar list lengths
        nmax = int(counts[order[min(i + 63, order.numel() - 1)]])
        j = i
        while j < order.numel() and (j - i) < 64:
            nj = int(counts[order[j]])
            if (j - i + 1) * 256 * max(nj, 1) > max_elems and j > i:
                break
            nmax = nj
            j += 1
        tiles_c = order[i:j]
        i = j
        nt = tiles_c.numel()
        cnt = counts[tiles_c]
        k = torch.arange(nmax, device=dev)
        in_list = k[None, :] < cnt[:, None]  # [nt, n]
        gpos = (tr[tiles_c, 0][:, None] + k[None, :]).clamp(max=max(idx_sorted.numel() - 1, 0))
        g = idx_sorted[gpos]  # [nt, n]
        tx = (tiles_c % gx).to(torch.float32) * BLOCK
        ty = (tiles_c // gx).to(torch.float32) * BLOCK
        pxf = tx[:, None] + lx[None, :]  # [nt, 256]
        pyf = ty[:, None] + ly[None, :]
        dx = uv[g][:, None, :, 0] - pxf[:, :, None]  # [nt,256,n]
        dy = uv[g][:, None, :, 1] - pyf[:, :, None]
        cn = conic[g]  # [nt,n,3]
        power = -0.5 * (cn[:, None, :, 0] * dx * dx + cn[:, None, :, 2] * dy * dy) - cn[:, None, :, 1] * dx * dy
        G = torch.exp(power)
        alpha_raw = opacity.reshape(-1)[g][:, None, :] * G
        # min(0.99, .) is applied to the value only: the reference backward differentiates
        # op*G as if unclamped (alpha_blending.cu:200-201,228 -- dL_dG = opac * dL_dalpha)
        alpha = alpha_raw + (torch.clamp(alpha_raw, max=0.99) - alpha_raw).detach()
        valid = in_list[:, None, :] & ~(power > 0) & ~(alpha < 1.0 / 255.0)
        a_eff = torch.where(valid, alpha, torch.zeros_like(alpha))
        T_incl = torch.cumprod(1.0 - a_eff, dim=-1)
        alive = valid & ~(T_incl < 1e-4)
        # once terminated, always terminated (T_incl is non-increasing)
        dead = torch.cummax((valid & (T_incl < 1e-4)).to(torch.int8), dim=-1).values.bool()
        alive = alive & ~dead
        a_use = torch.where(alive, alpha, torch.zeros_like(alpha))
        T_incl2 = torch.cumprod(1.0 - a_use, dim=-1)
        T_excl = torch.cat([torch.ones_like(T_incl2[..., :1]), T_incl2[..., :-1]], -1)
        wgt = a_use * T_excl  # [nt,256,n]
        Fpix = torch.einsum("tpn,tnc->tpc", wgt, feature[g])
        Tf = T_incl2[..., -1]
        last = torch.where(alive, (k + 1)[None, None, :].expand_as(alive), torch.zeros_like(alive, dtype=torch.int64)).amax(-1)
        piece = Fpix + Tf[..., None] * bg
        if dL_pad is not None:
            tyc = (tiles_c // gx)[:, None] * BLOCK + ly.long()[None, :]
            txc = (tiles_c % gx)[:, None] * BLOCK + lx.long()[None, :]
            if piece.requires_grad:
                torch.autograd.backward(piece, dL_pad[tyc, txc])
            piece, Tf = piece.detach(), Tf.detach()
        pieces_out.append(piece)
        pieces_T.append(Tf)
        pieces_n.append(last.to(torch.int32))
        pieces_tile.append(tiles_c)
    if pieces_tile:
        tiles_all = torch.cat(pieces_tile)
        o = torch.cat(pieces_out)  # [nt,256,C]
        tyi = (tiles_all // gx)[:, None] * BLOCK + ly.long()[None, :]
        txi = (tiles_all % gx)[:, None] * BLOCK + lx.long()[None, :]
        out = out.index_put((tyi, txi), o)
        final_T = final_T.index_put((tyi, txi), torch.cat(pieces_T))
        ncontrib = ncontrib.index_put((tyi, txi), torch.cat(pieces_n))
    img = out[:H, :W].permute(2, 0, 1).contiguous()
    if return_aux:
        return img, final_T[:H, :W].contiguous(), ncontrib[:H, :W].contiguous()
    return img


# ----------------------------------------------------------------------------
# a13 rasterization            msplat/msplat/__init__.py:22-93
# ----------------------------------------------------------------------------
def rasterization(xyz, scale, rotate, opacity, feature, intr, extr, W, H, bg, ndc=None):
    uv, depth = project_point(xyz, intr, extr, W, H)
    visible = depth != 0
    cov3d = compute_cov3d(scale, rotate, visible)
    conic, radius, tiles = ewa_project(xyz, cov3d, intr, extr, uv, W, H, visible)
    idx_sorted, tile_range = sort_gaussian(uv, depth, W, H, radius, tiles)
    return alpha_blending(uv, conic, opacity, feature, idx_sorted, tile_range, bg, W, H, ndc)


# ----------------------------------------------------------------------------
# a1  MsplatRender.render_iter  pointrix/model/renderer/msplat.py:94-158
# ----------------------------------------------------------------------------
def render_iter(
    height: int, width: int, extrinsic_matrix, intrinsic_params, camera_center,
    position, opacity, scaling, rotation, shs, sh_degree: int = 3, bg_color: float = 1.0,
    render_depth: bool = False, extra_features: Optional[Dict[str, torch.Tensor]] = None,
    tile_mask: Optional[torch.Tensor] = None,
) -> Dict:
    direction = position - camera_center.reshape(1, 3)
    direction = direction / direction.norm(dim=1, keepdim=True)
    sh_coeff = shs.permute(0, 2, 1)
    sh_mask = torch.zeros_like(sh_coeff)
    sh_mask[..., : (sh_degree + 1) ** 2] = 1.0
    rgb = compute_sh(sh_coeff * sh_mask, direction)
    rgb = (rgb + 0.5).clamp(min=0.0)
    extr = extrinsic_matrix[:3, :]
    uv, depth = project_point(position, intrinsic_params, extr, width, height, nearest=0.2)
    visible = depth != 0
    cov3d = compute_cov3d(scaling, rotation, visible)
    conic, radius, tiles = ewa_project(position, cov3d, intrinsic_params, extr, uv, width, height, visible)
    idx_sorted, tile_range = sort_gaussian(uv, depth, width, height, radius, tiles)
    feats = {"rgb": rgb}
    if render_depth:
        feats["depth"] = depth
    if extra_features:
        feats.update(extra_features)
    feature = torch.cat(list(feats.values()), dim=-1)
    ndc = torch.zeros_like(uv, requires_grad=True)
    img = alpha_blending(uv, conic, opacity, feature, idx_sorted, tile_range, bg_color, width, height, ndc,
                         tile_mask=tile_mask)
    split, s = {}, 0
    for k, v in feats.items():
        split[k] = img[s : s + v.shape[-1]]
        s += v.shape[-1]
    return {
        "rendered_features_split": split, "uv_points": ndc, "visibility": radius > 0, "radii": radius,
        "_aux": {"uv": uv, "depth": depth, "conic": conic, "tiles": tiles, "idx_sorted": idx_sorted,
                 "tile_range": tile_range, "rgb": rgb},
    }


def render_step(height: int, width: int, extrinsic_matrix, intrinsic_params, camera_center,
                position, opacity, scaling, rotation, shs, loss_fn, sh_degree: int = 3, bg_color: float = 1.0,
                tile_mask: Optional[torch.Tensor] = None) -> Dict:
    """One training iteration of :func:`render_iter` + ``loss_fn(rgb[3,H,W]) -> scalar`` + backward with
    bounded memory: the same math and the same gradients as ``loss_fn(render_iter(...)).backward()``, but
    the blend is differentiated chunk by chunk (``alpha_blending(dL_dout=...)``) behind a detached
    boundary, so full-size scenes fit the host.  Gradients land in ``.grad`` of the leaves passed in
    (Gaussian parameters and, when they require grad, the camera tensors); returns
    ``{"rgb", "loss", "radii", "ndc_grad", "isect_total", "isect_blended"}``.  ``tile_mask``: see
    :func:`alpha_blending` (a sample of the view's tiles; everything per-Gaussian still runs in full).
    """
    direction = position - camera_center.reshape(1, 3)
    direction = direction / direction.norm(dim=1, keepdim=True)
    sh_coeff = shs.permute(0, 2, 1)
    sh_mask = torch.zeros_like(sh_coeff)
    sh_mask[..., : (sh_degree + 1) ** 2] = 1.0
    rgb = (compute_sh(sh_coeff * sh_mask, direction) + 0.5).clamp(min=0.0)
    extr = extrinsic_matrix[:3, :]
    uv, depth = project_point(position, intrinsic_params, extr, width, height, nearest=0.2)
    visible = depth != 0
    cov3d = compute_cov3d(scaling, rotation, visible)
    conic, radius, tiles = ewa_project(position, cov3d, intrinsic_params, extr, uv, width, height, visible)
    idx_sorted, tile_range = sort_gaussian(uv, depth, width, height, radius, tiles)
    # detached boundary: the blend sees leaves of its own
    b_uv, b_conic, b_op, b_feat = (t.detach().requires_grad_() for t in (uv, conic, opacity, rgb))
    with torch.no_grad():
        img = alpha_blending(b_uv, b_conic, b_op, b_feat, idx_sorted, tile_range, bg_color, width, height,
                             tile_mask=tile_mask)
    img.requires_grad_()
    loss = loss_fn(img)
    loss.backward()
    alpha_blending(b_uv, b_conic, b_op, b_feat, idx_sorted, tile_range, bg_color, width, height, dL_dout=img.grad,
                   tile_mask=tile_mask)
    zero = torch.zeros_like
    g = [b_uv.grad if b_uv.grad is not None else zero(b_uv), b_conic.grad if b_conic.grad is not None else zero(b_conic),
         b_op.grad if b_op.grad is not None else zero(b_op), b_feat.grad if b_feat.grad is not None else zero(b_feat)]
    # dL/duv reaches nothing upstream (ewa_project.py:84-85 returns None for uv; project_point gets it): feed
    # the four boundary gradients back into the per-Gaussian graph
    torch.autograd.backward([uv, conic, opacity, rgb], g)
    scale = torch.tensor([0.5 * width, 0.5 * height], dtype=uv.dtype)
    per_tile = (tile_range[:, 1] - tile_range[:, 0]).to(torch.int64)
    return {"rgb": img.detach(), "loss": loss.detach(), "radii": radius, "ndc_grad": g[0] * scale,
            "isect_total": int(per_tile.sum()),
            "isect_blended": int(per_tile.sum() if tile_mask is None else per_tile[tile_mask].sum())}
