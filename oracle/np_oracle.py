"""Second, independent restatement of simulator.rs in NumPy float32.

TEST INFRASTRUCTURE ONLY.  Written separately from fluid_oracle.c (different structure:
vectorised phases, projection swept by anti-diagonal wavefronts) so that a transcription slip
in either is caught by the other: tests require the two to agree bit-for-bit, and that
agreement is the source of tests/golden/.  PARITY UNPINNED against the Rust binary (no Rust
toolchain here, the reference has no tests) -- see fluid_oracle.c's header.

All arithmetic is done on np.float32 arrays/scalars so every operation rounds once to f32,
in the association order of the Rust source (cited per line).
"""
from __future__ import annotations

import numpy as np

F = np.float32


def _as_usize(x):
    """Rust `f32 as usize` (saturating, NaN -> 0) for values far below 2^63."""
    x = np.asarray(x, dtype=np.float32)
    return np.where(x > 0, x, F(0)).astype(np.int64)


class NpSim:
    def __init__(self, width, height, gravity=0.0, wind_speed=50.0, smoke_size=0.25, density=1000.0,
                 iterations=50, omega=1.9):
        assert width >= 1 and height >= 1
        self.gravity, self.wind_speed = F(gravity), F(wind_speed)
        self.smoke_size, self.density = F(smoke_size), F(density)
        self.iterations, self.omega = iterations, F(omega)
        self.W, self.H = width, height
        self.m = np.zeros(width * height, np.uint8)
        self.m[:height] = 1                         # simulator.rs:113-119
        self._reset_fields()

    # simulator.rs:73-111
    def _reset_fields(self):
        n = self.W * self.H
        self.u = np.zeros(n, F)
        self.v = np.zeros(n, F)
        self.p = np.zeros(n, F)
        self.s = np.ones(n, F)
        self._wind_and_pipe()

    def _wind_and_pipe(self):
        H, n = self.H, self.W * self.H
        self.u[H:min(2 * H, n)] = self.wind_speed   # :79-85
        self.s[:min(H, n)] = F(1.0)                 # :96-99
        pipe = F(H) * self.smoke_size               # :101
        mid = F(H) * F(0.5)                         # :102
        lo = int(_as_usize(mid - pipe * F(0.5)))    # :103
        hi = int(_as_usize(mid + pipe * F(0.5)))    # :104
        self.s[lo:min(hi, n)] = F(0.0)              # :106-110 (empty if lo >= hi)

    # simulator.rs:43-53, 121-135
    def resize(self, width, height):
        old_h = self.H
        self.m[:old_h] = 0
        m = np.zeros(width * height, np.uint8)
        k = min(m.size, self.m.size)
        m[:k] = self.m[:k]
        m[:height] = 1
        self.m = m
        self.W, self.H = width, height
        self._reset_fields()

    def restart_sim(self):
        self.resize(self.W, self.H)

    # simulator.rs:67-71
    def set_config(self, gravity, wind_speed, smoke_size, density):
        self.wind_speed, self.smoke_size = F(wind_speed), F(smoke_size)
        self._wind_and_pipe()
        self.gravity, self.density = F(gravity), F(density)

    def set_block(self, x, y, value=1):
        self.m[x * self.H + y] = value              # :331-340

    # simulator.rs:365-375
    def _border(self):
        W, H = self.W, self.H
        idx = np.arange(W * H)
        rem = idx % H
        return (idx < H) | (idx >= (W - 1) * H) | (rem == H - 1) | (rem == 0)

    def _skip(self):
        return (self.m != 0) | self._border()

    # simulator.rs:137-149
    def add_gravity(self, dt):
        dt = F(dt)
        live = ~self._skip()
        self.v[live] = self.v[live] + (self.gravity * dt)

    # simulator.rs:151-190, swept by wavefronts t = i + j + 2k.  Every (i, j, k) on one
    # wavefront touches locations no other member touches, and per location the order of
    # touches equals the lexicographic sweep's, so the result is identical.
    def make_incompressible(self, dt, lexicographic=False):
        dt = F(dt)
        W, H, K = self.W, self.H, self.iterations
        self.p[:] = F(0.0)                                   # :154
        pc = self.density / dt                               # :155
        if lexicographic:
            return self._project_lex(pc)
        skip = self._skip()
        u, v, p, m = self.u, self.v, self.p, self.m
        ii, jj = np.meshgrid(np.arange(W), np.arange(H), indexing="ij")
        diag = (ii + jj).ravel()
        order = np.argsort(diag, kind="stable")
        starts = np.searchsorted(diag[order], np.arange(W + H))
        cells_on = [order[starts[d]:starts[d + 1]] if d + 1 < W + H else order[starts[d]:]
                    for d in range(W + H - 1)]
        live_on = [c[~skip[c]] for c in cells_on]
        for t in range(W + H - 1 + 2 * (K - 1)):
            for k in range(K):
                d = t - 2 * k
                if d < 0 or d >= W + H - 1:
                    continue
                idx = live_on[d]
                if idx.size == 0:
                    continue
                top, right, bottom, left = idx + 1, idx + H, idx - 1, idx - H   # :343-350
                sT = (m[top] == 0).astype(F)                                   # :165-168
                sR = (m[right] == 0).astype(F)
                sB = (m[bottom] == 0).astype(F)
                sL = (m[left] == 0).astype(F)
                nfl = ((sR + sT) + sL) + sB                                    # :169-170
                ok = nfl != 0                                                  # :172
                idx, top, right = idx[ok], top[ok], right[ok]
                sT, sR, sB, sL, nfl = sT[ok], sR[ok], sB[ok], sL[ok], nfl[ok]
                div = ((u[right] - u[idx]) + v[top]) - v[idx]                  # :176-178
                corr = self.omega * ((-div) / nfl)                             # :180
                u[idx] = u[idx] - corr * sL                                    # :181
                u[right] = u[right] + corr * sR                                # :182
                v[idx] = v[idx] - corr * sB                                    # :184
                v[top] = v[top] + corr * sT                                    # :185
                p[idx] = p[idx] + pc * corr                                    # :186

    def _project_lex(self, pc):
        """Literal triple loop (simulator.rs:157-189); slow, for tiny grids only."""
        W, H = self.W, self.H
        skip = self._skip()
        u, v, p, m = self.u, self.v, self.p, self.m
        for _ in range(self.iterations):
            for i in range(W):
                for j in range(H):
                    c = i * H + j
                    if skip[c]:
                        continue
                    t, r, b, l = c + 1, c + H, c - 1, c - H
                    sT, sR = F(m[t] == 0), F(m[r] == 0)
                    sB, sL = F(m[b] == 0), F(m[l] == 0)
                    nfl = ((sR + sT) + sL) + sB
                    if nfl == 0:
                        continue
                    div = ((u[r] - u[c]) + v[t]) - v[c]
                    corr = self.omega * ((-div) / nfl)
                    u[c] = u[c] - corr * sL
                    u[r] = u[r] + corr * sR
                    v[c] = v[c] - corr * sB
                    v[t] = v[t] + corr * sT
                    p[c] = p[c] + pc * corr

    # simulator.rs:269-298
    def _sample(self, x, y, f, dx, dy):
        W, H = self.W, self.H
        x = np.fmax(np.fmin(x, F(W)), F(1.0))                # :271 (min/max ignore NaN)
        y = np.fmax(np.fmin(y, F(H)), F(1.0))                # :272
        xs = x - F(dx)
        ys = y - F(dy)
        xl = np.minimum(_as_usize(np.floor(xs)), W - 1)      # :283
        xr = np.minimum(xl + 1, W - 1)                       # :284
        tx = xs - xl.astype(F)                               # :285
        yb = np.minimum(_as_usize(np.trunc(ys)), H - 1)      # :287
        yt = np.minimum(yb + 1, H - 1)                       # :288
        ty = ys - yb.astype(F)                               # :289
        sx = F(1.0) - tx
        sy = F(1.0) - ty
        return ((((sx * sy) * f[xl * H + yb]) + ((tx * sy) * f[xr * H + yb]))
                + ((tx * ty) * f[xr * H + yt])) + ((sx * ty) * f[xl * H + yt])   # :294-297

    # simulator.rs:192-267
    def move_velocity(self, dt):
        dt = F(dt)
        W, H = self.W, self.H
        u, v, s = self.u, self.v, self.s
        live = np.nonzero(~self._skip())[0]
        i, j = live // H, live % H
        fi, fj = i.astype(F), j.astype(F)
        half = F(0.5)
        # u component, :207-213 and :245-255 (sum starts at 0)
        avg_v = ((((F(0) + v[(i - 1) * H + j]) + v[(i - 1) * H + j + 1]) + v[i * H + j + 1])
                 + v[i * H + j]) * F(0.25)
        nu_live = self._sample(fi - u[live] * dt, (fj + half) - avg_v * dt, u, 0.0, 0.5)
        # v component, :216-224 and :257-267
        avg_u = ((((F(0) + u[i * H + j - 1]) + u[(i + 1) * H + j - 1]) + u[(i + 1) * H + j])
                 + u[i * H + j]) * F(0.25)
        nv_live = self._sample((fi + half) - avg_u * dt, fj - v[live] * dt, v, 0.5, 0.0)
        # smoke, :227-237
        cv = (v[live] + v[i * H + j + 1]) * half
        cu = (u[live] + u[(i + 1) * H + j]) * half
        ns_live = self._sample((fi + half) - cu * dt, (fj + half) - cv * dt, s, 0.5, 0.5)
        nu, nv, ns = u.copy(), v.copy(), s.copy()
        nu[live], nv[live], ns[live] = nu_live, nv_live, ns_live
        self.u, self.v, self.s = nu, nv, ns

    # simulator.rs:59-65
    def step(self, dt, lexicographic=False):
        self.add_gravity(dt)
        self.make_incompressible(dt, lexicographic=lexicographic)
        self.move_velocity(dt)


def _as_u8(x):
    """Rust `f32 as u8` (saturating, NaN -> 0), vectorised"""
    x = np.asarray(x, np.float32)
    with np.errstate(invalid="ignore"):
        y = np.where(x > 0, np.minimum(x, F(255.0)), F(0.0))
    return np.trunc(y).astype(np.uint8)


def render_view(p, s, m, W, H, area_w, area_h):
    """Independent restatement of render_sim / get_color / get_linear_gradient
    (src/ui/sim_renderer.rs:23-151) on float32 arrays; returns (bytes[area_h, area_w, 2, 4], min, max)."""
    p = np.asarray(p, F); s = np.asarray(s, F)
    mx, mn = F(-3.402823466e38), F(3.402823466e38)
    for v in p:                                   # the literal `if < min {..} else if > max {..}` scan (:31-37)
        if v < mn:
            mn = v
        elif v > mx:
            mx = v
    out = np.zeros((area_h, area_w, 2, 4), np.uint8)
    rows = np.arange(area_h)
    for half, yy in ((0, H - 1 - 2 * rows), (1, H - 2 - 2 * rows)):
        idx = (np.arange(area_w)[None, :] * H + yy[:, None])
        val, smoke, blk = p[idx], s[idx], m[idx] != 0
        with np.errstate(all="ignore"):
            diff = mx - mn
            norm = np.full(val.shape, F(0.5), F) if diff == 0 else ((val - mn) / diff).astype(F)
            ctype = np.floor(norm * F(4.0)).astype(F)
            cval = ((norm - ctype * F(0.25)) / F(0.25)).astype(F)
            ct = _as_u8(ctype)
            one, zero = np.ones_like(cval), np.zeros_like(cval)
            r = np.select([ct == 0, ct == 1, ct == 2, ct == 3], [zero, zero, cval, one], one)
            g = np.select([ct == 0, ct == 1, ct == 2, ct == 3], [cval, one, one, one - cval], zero)
            b = np.select([ct == 0, ct == 1, ct == 2, ct == 3], [one, one - cval, zero, zero], zero)
            rgb = np.stack([_as_u8(F(255.0) * r), _as_u8(F(255.0) * g), _as_u8(F(255.0) * b)], -1).astype(np.int32)
            red = _as_u8(F(255.0) * smoke).astype(np.int32)[..., None]
        col = np.clip(rgb - red, 0, 255).astype(np.uint8)
        out[:, :, half, :3] = np.where(blk[..., None], 255, col)
        out[:, :, half, 3] = blk
    return out, float(mn), float(mx)
