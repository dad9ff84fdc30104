"""Independent witness for the light-sampling half of the path (TEST INFRASTRUCTURE, like everything under oracle/).

Written from the reference's shader sources only — NOT from oracle/vr_oracle.cpp or the CUDA kernels — in numpy float32:

  * env-map importance map      F/Experimental/Scene/Lights/EnvMapSamplerSetup.cs.slang:48-75 (512^2 texels, 8x8 sub-samples,
                                octahedral equal-area map -> lat-long lookup -> luminance) + its mip chain
  * hierarchical env sampling   F/Experimental/Scene/Lights/EnvMapSampler.slang:92-165 (2x2 warps from the 1x1 mip down)
  * env-map evaluation          F/Experimental/Scene/Lights/EnvMap.slang (toLocal, lat-long, bilinear, intensity * tint)
  * octahedral / lat-long maps  F/Utils/Math/MathHelpers.slang:92-99,200-219
  * Henyey-Greenstein phase function and its sampling   VR/VolumeBase.slang:50-96

Texture filtering is a hardware behaviour the shader source does not spell out; this file uses the D3D rule (texel centres at
integer + 0.5, bilinear weights from the fractional part, wrap in U, clamp in V), which DESIGN.md section 2 pins.
tests/test_light_witness.py holds the C++ oracle against this file."""
import numpy as np

F = np.float32
PI = F(3.14159265358979323846)


def luminance(rgb):
    return F(0.2126) * rgb[..., 0] + F(0.7152) * rgb[..., 1] + F(0.0722) * rgb[..., 2]


def oct_to_ndir_equal_area_unorm(p):
    """MathHelpers.slang:200-219; p: (..., 2) float32 in [0,1)^2 -> unit vectors (..., 3)."""
    p = p.astype(F) * F(2) - F(1)
    ax, ay = np.abs(p[..., 0]), np.abs(p[..., 1])
    d = F(1) - (ax + ay)
    r = F(1) - np.abs(d)
    with np.errstate(divide="ignore", invalid="ignore"):
        phi = np.where(r > 0, ((ay - ax) / r + F(1)) * (PI / F(4)), F(0)).astype(F)
    f = (r * np.sqrt(F(2) - r * r)).astype(F)
    x = f * np.sign(p[..., 0]) * np.cos(phi)
    y = f * np.sign(p[..., 1]) * np.sin(phi)
    z = np.sign(d) * (F(1) - r * r)
    return np.stack([x, y, z], axis=-1).astype(F)


def world_to_latlong_map(d):
    """MathHelpers.slang:92-99."""
    n = d / np.sqrt(np.sum(d * d, axis=-1, keepdims=True, dtype=F)).astype(F)
    u = np.arctan2(n[..., 0], -n[..., 2]).astype(F) * (F(1) / (F(2) * PI)) + F(0.5)
    v = np.arccos(np.clip(n[..., 1], -1, 1)).astype(F) * (F(1) / PI)
    return np.stack([u, v], axis=-1).astype(F)


def bilinear(tex, uv):
    """tex: (H, W, C) float32; uv: (..., 2).  Wrap in U, clamp in V, texel centres at +0.5."""
    h, w = tex.shape[:2]
    x = uv[..., 0].astype(F) * F(w) - F(0.5)
    y = uv[..., 1].astype(F) * F(h) - F(0.5)
    x0 = np.floor(x); y0 = np.floor(y)
    fx = (x - x0).astype(F)[..., None]; fy = (y - y0).astype(F)[..., None]
    x0 = x0.astype(np.int64); y0 = y0.astype(np.int64)
    xa, xb = np.mod(x0, w), np.mod(x0 + 1, w)
    ya, yb = np.clip(y0, 0, h - 1), np.clip(y0 + 1, 0, h - 1)
    top = tex[ya, xa] + fx * (tex[ya, xb] - tex[ya, xa])
    bot = tex[yb, xa] + fx * (tex[yb, xb] - tex[yb, xa])
    return (top + fy * (bot - top)).astype(F)


def importance_mips(texels, dim=512, spp=64):
    """EnvMapSamplerSetup.cs.slang:48-75 + the mip chain (each level the mean of its 2x2 children); finest first."""
    sx = max(1, int(np.sqrt(spp))); sy = spp // sx
    rgb = np.ascontiguousarray(texels[..., :3], dtype=F)
    base = np.zeros((dim, dim), dtype=F)
    py, px = np.meshgrid(np.arange(dim), np.arange(dim), indexing="ij")
    for y in range(sy):
        for x in range(sx):
            pos = np.stack([(px * sx + x + 0.5) / (dim * sx), (py * sy + y + 0.5) / (dim * sy)], axis=-1).astype(F)
            base += luminance(bilinear(rgb, world_to_latlong_map(oct_to_ndir_equal_area_unorm(pos))))
    base = (base * F(1.0 / (sx * sy))).astype(F)
    mips = [base]
    while mips[-1].shape[0] > 1:
        m = mips[-1]
        mips.append(((m[0::2, 0::2] + m[0::2, 1::2] + m[1::2, 0::2] + m[1::2, 1::2]) * F(0.25)).astype(F))
    return mips


def env_sample(mips, u0, u1):
    """EnvMapSampler.slang:92-165 (local frame): returns (direction, pdf w.r.t. solid angle, chosen texel)."""
    p = [F(u0), F(u1)]
    pos = [0, 0]
    base_mip = len(mips) - 1
    for mip in range(base_mip - 1, -1, -1):
        pos = [pos[0] * 2, pos[1] * 2]
        m = mips[mip]
        w = [m[pos[1], pos[0]], m[pos[1], pos[0] + 1], m[pos[1] + 1, pos[0]], m[pos[1] + 1, pos[0] + 1]]
        q = [F(w[0] + w[2]), F(w[1] + w[3])]
        d = F(q[0] / F(q[0] + q[1]))
        if p[0] < d:
            ox = 0; p[0] = F(p[0] / d)
        else:
            ox = 1; p[0] = F(F(p[0] - d) / F(F(1) - d))
        e = F(w[ox] / q[ox])
        if p[1] < e:
            oy = 0; p[1] = F(p[1] / e)
        else:
            oy = 1; p[1] = F(F(p[1] - e) / F(F(1) - e))
        pos = [pos[0] + ox, pos[1] + oy]
    dim = mips[0].shape[0]
    uv = np.array([F(F(pos[0]) + p[0]) * F(1.0 / dim), F(F(pos[1]) + p[1]) * F(1.0 / dim)], dtype=F)
    direction = oct_to_ndir_equal_area_unorm(uv)
    pdf = F(mips[0][pos[1], pos[0]] / mips[base_mip][0, 0]) * F(1.0 / (4.0 * np.pi))
    return direction, float(pdf), tuple(pos)


def env_eval(texels, direction, intensity=1.0, tint=(1.0, 1.0, 1.0)):
    rgb = np.ascontiguousarray(texels[..., :3], dtype=F)
    return (F(intensity) * np.asarray(tint, dtype=F) * bilinear(rgb, world_to_latlong_map(np.asarray(direction, dtype=F)))).astype(F)


def phase_hg(cos_theta, g):
    """VR/VolumeBase.slang:64-68."""
    cos_theta, g = F(cos_theta), F(g)
    denom = F(1) + g * g + F(2) * g * cos_theta
    return float(F(0.07957747154594766788) * (F(1) - g * g) / (denom * np.sqrt(denom)))


def sample_phase(g, wo, u0, u1):
    """VR/VolumeBase.slang:50-62,81-96: returns (wi, pdf)."""
    g, u0, u1 = F(g), F(u0), F(u1)
    wo = np.asarray(wo, dtype=F)
    if abs(g) < F(1e-3):
        cos_t = F(1) - F(2) * u0
    else:
        sqr = (F(1) - g * g) / (F(1) + g - F(2) * g * u0)
        cos_t = -(F(1) + g * g - sqr * sqr) / (F(2) * g)
    sin_t = np.sqrt(max(F(0), F(1) - cos_t * cos_t)).astype(F)
    phi = F(2) * PI * u1
    if abs(wo[0]) > abs(wo[1]):
        v2 = np.array([-wo[2], 0, wo[0]], dtype=F) / np.sqrt(wo[0] * wo[0] + wo[2] * wo[2]).astype(F)
    else:
        v2 = np.array([0, wo[2], -wo[1]], dtype=F) / np.sqrt(wo[1] * wo[1] + wo[2] * wo[2]).astype(F)
    v3 = np.cross(wo, v2).astype(F)
    wi = (sin_t * np.cos(phi).astype(F) * v2 + sin_t * np.sin(phi).astype(F) * v3 + cos_t * wo).astype(F)
    return wi, phase_hg(cos_t, g)
