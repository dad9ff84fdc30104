"""TEST INFRASTRUCTURE (oracle): numpy restatement of the reference's AccumulatePass and ErrorMeasurePass arithmetic.
Only tests/ may import this.  Pinned by hand-computed known-answer cases in tests/test_post_oracle.py (the reference holds
no fixtures for these passes).

  accumulate_*   Source/RenderPasses/AccumulatePass/Accumulate.cs.slang:57-122 (+ the clear at frame 0, AccumulatePass.cpp)
  error_measure  Source/RenderPasses/ErrorMeasurePass/ErrorMeasurer.cs.slang:41-61, ErrorMeasurePass.cpp:241-243
"""
import numpy as np

F = np.float32


class Accumulator:
    def __init__(self, mode="Double"):
        self.mode = mode
        self.count = 0
        self.sum = None
        self.corr = None

    def add(self, cur):
        cur = np.asarray(cur, dtype=F)
        if self.count == 0:
            self.sum = np.zeros(cur.shape, dtype=np.float64 if self.mode == "Double" else F)
            self.corr = np.zeros(cur.shape, dtype=F)
        n = self.count + 1
        if self.mode == "Single":                      # :57-70
            self.sum = (self.sum + cur).astype(F)
            out = (self.sum / F(n)).astype(F)
        elif self.mode == "SingleCompensated":         # :74-93
            y = (cur - self.corr).astype(F)
            nxt = (self.sum + y).astype(F)
            out = (nxt / F(n)).astype(F)
            self.corr = ((nxt - self.sum).astype(F) - y).astype(F)
            self.sum = nxt
        else:                                          # :97-122
            self.sum = self.sum + cur.astype(np.float64)
            out = (self.sum / np.float64(n)).astype(F)
        self.count += 1
        return out


def error_measure(source, reference, world_position=None, ignore_background=True, squared=True, average=False):
    s = np.asarray(source, dtype=F)[..., :3]
    r = np.asarray(reference, dtype=F)[..., :3]
    d = np.abs((s - r).astype(F))
    if ignore_background and world_position is not None:
        d = np.where((np.asarray(world_position)[..., 3] != 0)[..., None], d, F(0))
    if squared:
        d = (d * d).astype(F)
    if average:
        a = (((d[..., 0] + d[..., 1]).astype(F) + d[..., 2]).astype(F) / F(3)).astype(F)
        d = np.stack([a, a, a], axis=-1)
    total = d.reshape(-1, 3).astype(np.float64).sum(axis=0)
    npx = F(s.shape[0] * s.shape[1])
    err = (total.astype(F) / npx).astype(F)
    avg = (((err[0] + err[1]).astype(F) + err[2]).astype(F) / F(3)).astype(F)
    return d, err, avg


# ---- ToneMapper (Source/RenderPasses/ToneMapper/{ToneMapping,Luminance}.ps.slang, ToneMapper.cpp:502-522, ColorUtils.h:60-216) ----
F = np.float32
_RGB2XYZ = np.array([[0.4123907992659595, 0.3575843393838780, 0.1804807884018343], [0.2126390058715104, 0.7151686787677559, 0.0721923153607337],
                     [0.0193308187155918, 0.1191947797946259, 0.9505321522496608]])
_XYZ2RGB = np.array([[3.2409699419045213, -1.5373831775700935, -0.4986107602930033], [-0.9692436362808798, 1.8759675015077206, 0.0415550574071756],
                     [0.0556300796969936, -0.2039769588889765, 1.0569715142428784]])
_XYZ2LMS = np.array([[0.7328, 0.4296, -0.1624], [-0.7036, 1.6975, 0.0061], [0.0030, 0.0136, 0.9834]])
_LMS2XYZ = np.array([[1.096123820835514, -0.278869000218287, 0.182745179382773], [0.454369041975359, 0.473533154307412, 0.072097803717229],
                     [-0.009627608738429, -0.005698031216113, 1.015325639954543]])


def _temperature_to_xyz(T):
    t = float(T)
    xc = (-0.2661239e9 / t ** 3 - 0.2343580e6 / t ** 2 + 0.8776956e3 / t + 0.179910) if T < 4000 else (-3.0258469e9 / t ** 3 + 2.1070379e6 / t ** 2 + 0.2226347e3 / t + 0.240390)
    x = xc
    if T < 2222:
        yc = -1.1063814 * x ** 3 - 1.34811020 * x ** 2 + 2.18555832 * x - 0.20219683
    elif T < 4000:
        yc = -0.9549476 * x ** 3 - 1.37418593 * x ** 2 + 2.09137015 * x - 0.16748867
    else:
        yc = 3.0817580 * x ** 3 - 5.87338670 * x ** 2 + 3.75112997 * x - 0.37001483
    return np.array([xc / yc, 1.0, (1.0 - xc - yc) / yc])


def tonemap_color_transform(exposureCompensation=0.0, autoExposure=False, filmSpeed=100.0, whiteBalance=False, whitePoint=6500.0, fNumber=1.0, shutter=1.0):
    """3x3 matrix M with c' = M c (ToneMapper::updateColorTransform): white balance x 2^compensation x manual exposure scale."""
    wb = np.eye(3)
    if whiteBalance:
        scale = (_XYZ2LMS @ _temperature_to_xyz(6500.0)) / (_XYZ2LMS @ _temperature_to_xyz(whitePoint))
        wb = (_XYZ2RGB @ _LMS2XYZ) @ np.diag(scale) @ (_XYZ2LMS @ _RGB2XYZ)
    manual = 1.0 if autoExposure else (0.01 * filmSpeed) / (shutter * fNumber * fNumber)
    return wb * (2.0 ** exposureCompensation) * manual


def tonemap_avg_log_luminance(img):
    """Luminance pass into the lower-power-of-two target (bilinear, wrap) + the 2x2 box mip chain down to 1x1."""
    h, w = img.shape[:2]
    w2 = 1 << (w.bit_length() - 1)
    h2 = 1 << (h.bit_length() - 1)
    u = (np.arange(w2, dtype=F) + F(0.5)) / F(w2) * F(w) - F(0.5)
    v = (np.arange(h2, dtype=F) + F(0.5)) / F(h2) * F(h) - F(0.5)
    u0, v0 = np.floor(u), np.floor(v)
    fu, fv = (u - u0).astype(F), (v - v0).astype(F)
    x0 = u0.astype(np.int64) % w; x1 = (x0 + 1) % w
    y0 = v0.astype(np.int64) % h; y1 = (y0 + 1) % h
    rgb = img[..., :3].astype(F)
    fma = lambda t, d, a: (t.astype(np.float64) * d.astype(np.float64) + a.astype(np.float64)).astype(F)
    a, b = rgb[y0][:, x0], rgb[y0][:, x1]
    c, d = rgb[y1][:, x0], rgb[y1][:, x1]
    fu3, fv3 = fu[None, :, None], fv[:, None, None]
    top = fma(np.broadcast_to(fu3, a.shape), (b - a).astype(F), a)
    bot = fma(np.broadcast_to(fu3, a.shape), (d - c).astype(F), c)
    r = fma(np.broadcast_to(fv3, a.shape), (bot - top).astype(F), top)
    lum = ((r[..., 0] * F(0.299)).astype(F) + (r[..., 1] * F(0.587)).astype(F)).astype(F)
    lum = (lum + (r[..., 2] * F(0.114)).astype(F)).astype(F)
    cur = np.log2(np.maximum(F(0.0001), lum)).astype(F)
    while cur.shape[0] > 1 or cur.shape[1] > 1:
        ch, cw = cur.shape
        nh, nw = max(1, ch // 2), max(1, cw // 2)
        ys0, ys1 = np.minimum(2 * np.arange(nh), ch - 1), np.minimum(2 * np.arange(nh) + 1, ch - 1)
        xs0, xs1 = np.minimum(2 * np.arange(nw), cw - 1), np.minimum(2 * np.arange(nw) + 1, cw - 1)
        cur = (((cur[ys0][:, xs0] + cur[ys0][:, xs1]).astype(F) + (cur[ys1][:, xs0] + cur[ys1][:, xs1]).astype(F)).astype(F) * F(0.25)).astype(F)
    return float(cur[0, 0])


def _saturate(x):
    """HLSL saturate: clamp to [0, 1], NaN -> 0 (a black pixel gives 0/0 in the Reinhard operators)."""
    return np.where(np.isnan(x), 0.0, np.clip(x, 0.0, 1.0))


def tonemap(img, M, op="Aces", autoExposure=False, clamp=True, whiteScale=11.2, whiteMaxLuminance=1.0):
    c = img[..., :3].astype(np.float64)
    if autoExposure:
        c = c * (0.042 / 2.0 ** tonemap_avg_log_luminance(img))
    c = c @ np.asarray(M, dtype=np.float64).T
    lum = lambda x: x @ np.array([0.299, 0.587, 0.114])
    with np.errstate(divide="ignore", invalid="ignore"):
        if op == "Reinhard":
            l = lum(c); c = c * ((l / (l + 1)) / l)[..., None]
        elif op == "ReinhardModified":
            l = lum(c); c = c * ((l * (1 + l / whiteMaxLuminance ** 2) * (1 + l)) / l)[..., None]
        elif op == "HejiHableAlu":
            c = np.maximum(0.0, c - 0.004); c = (c * (6.2 * c + 0.5)) / (c * (6.2 * c + 1.7) + 0.06); c = c ** 2.2
        elif op == "HableUc2":
            A, B, Cc, D, E, Fq = 0.22, 0.3, 0.1, 0.2, 0.01, 0.3
            uc2 = lambda x: ((x * (A * x + Cc * B) + D * E) / (x * (A * x + B) + D * Fq)) - E / Fq
            c = uc2(2.0 * c) * (1.0 / uc2(max(0.001, whiteScale)))
        elif op == "Aces":
            c = c * 0.6; c = _saturate((c * (2.51 * c + 0.03)) / (c * (2.43 * c + 0.59) + 0.14))
    if clamp:
        c = _saturate(c)
    out = img.astype(np.float64).copy()
    out[..., :3] = c
    return out
