"""The four panels of the reference's plot4Figures.m -- charging current, cell voltage, SOC and side-reaction overpotential
against time, each with its limits drawn in (plot4Figures.m:27-75) -- written as a self-contained SVG file.

plot4Figures.m draws ONE cell with MATLAB graphics (out of this build's scope).  Here the same figure is
produced from what the CUDA path logged (CellBatch.logs()): one cell's traces, or for a population the median with the
5-95 % band and the extremes (mek/analytics.py::trajectory_envelopes).  Plain text output, no plotting library.
"""
from __future__ import annotations

import numpy as np

from .analytics import trajectory_envelopes

_PANELS = (("u", "Charging Current", "Current [A]"), ("V", "Voltage", "Voltage [V]"), ("SOC", "SOC", "SOC [-]"),
           ("phise", "Side-reaction Overpotential", "Overpot. [V]"))


def _ticks(lo, hi, n=5):
    if not np.isfinite(lo) or not np.isfinite(hi) or hi <= lo:
        return [lo]
    step = 10.0 ** np.floor(np.log10((hi - lo) / n))
    for m in (1.0, 2.0, 5.0, 10.0):
        if (hi - lo) / (step * m) <= n:
            step *= m
            break
    t0 = np.ceil(lo / step) * step
    return [float(t0 + i * step) for i in range(int((hi - t0) / step) + 1)]


def plot4figures_svg(logs, path, limits=None, ref=None, Ts=1.0, cell=None, title="MPC Fast-Charge (4x)"):
    """logs: what CellBatch.logs() returns ({'u','V','SOC','phise'}: [N, cells]).  cell: index of the single cell to draw
    (plot4Figures.m's own figure); None draws the population (median, 5-95 % band, min / max).  limits: dict with any of
    u_max, u_min, v_max, v_min, phise_min (mpcData.const, plot4Figures.m:30-31,41-42,68-70); ref: target SOC in percent
    (mpcData.ref, :51-53).  Returns the SVG text (also written to `path` when given)."""
    limits = dict(limits or {})
    N = np.asarray(logs["u"]).shape[0]
    t = np.arange(N) * float(Ts)
    if cell is None:
        env = trajectory_envelopes(logs, (0.0, 0.05, 0.5, 0.95, 1.0))
    W, H, L, R, T, B = 680, 820, 70, 20, 34, 30
    ph = (H - 40) / 4.0
    out = ['<svg xmlns="http://www.w3.org/2000/svg" width="%d" height="%d" viewBox="0 0 %d %d" font-family="sans-serif" font-size="11">' % (W, H, W, H),
           '<rect width="100%%" height="100%%" fill="white"/><text x="%d" y="16" font-size="13" text-anchor="middle">%s</text>' % (W // 2, title)]
    lines_of = {"u": [("u_max", "u_max"), ("u_min", "u_min")], "V": [("v_max", "v_max"), ("v_min", "v_min")], "SOC": [], "phise": [("phise_min", "phi_min")]}
    for p, (key, ttl, ylab) in enumerate(_PANELS):
        y0 = 20 + p * ph + T
        h = ph - T - B
        w = W - L - R
        if cell is None:
            series = env[key]
            ys = series[[0, 4]]
        else:
            series = np.asarray(logs[key], float)[:, cell][None, :]
            ys = series
        refs = [(float(limits[k]), lab) for k, lab in lines_of[key] if k in limits and np.isfinite(float(np.mean(limits[k])))]
        if key == "SOC" and ref is not None:
            refs.append((float(ref) / 100.0, "Target"))
        if key == "phise":
            refs.append((0.0, ""))
        vals = np.concatenate([np.ravel(ys[np.isfinite(ys)]), [r for r, _ in refs]]) if np.isfinite(ys).any() else np.array([0.0, 1.0])
        lo, hi = float(vals.min()), float(vals.max())
        if key == "SOC":
            lo, hi = 0.0, 1.0                                       # plot4Figures.m:57
        pad = 0.06 * (hi - lo if hi > lo else 1.0)
        lo, hi = lo - pad, hi + pad
        X = lambda tt: L + w * (np.asarray(tt) - t[0]) / max(t[-1] - t[0], 1e-12)
        Y = lambda v: y0 + h * (1.0 - (np.asarray(v) - lo) / (hi - lo))
        out.append('<g id="panel-%s"><rect x="%d" y="%.1f" width="%d" height="%.1f" fill="none" stroke="#444"/>' % (key, L, y0, w, h))
        out.append('<text x="%d" y="%.1f" font-size="12" text-anchor="middle">%s</text>' % (L + w // 2, y0 - 6, ttl))
        out.append('<text transform="translate(14 %.1f) rotate(-90)" text-anchor="middle">%s</text>' % (y0 + h / 2, ylab))
        for tk in _ticks(lo, hi):
            out.append('<line x1="%d" x2="%d" y1="%.1f" y2="%.1f" stroke="#ddd"/><text x="%d" y="%.1f" text-anchor="end">%g</text>' %
                       (L, L + w, Y(tk), Y(tk), L - 4, Y(tk) + 4, round(tk, 6)))
        for tk in _ticks(t[0], t[-1], 6):
            out.append('<line x1="%.1f" x2="%.1f" y1="%.1f" y2="%.1f" stroke="#ddd"/>' % (X(tk), X(tk), y0, y0 + h))
            if p == 3:
                out.append('<text x="%.1f" y="%.1f" text-anchor="middle">%g</text>' % (X(tk), y0 + h + 14, tk))

        def poly(v, **kw):
            ok = np.isfinite(v)
            pts = " ".join("%.1f,%.1f" % (x, y) for x, y in zip(X(t[ok]), Y(v[ok])))
            return '<polyline points="%s" fill="none" %s/>' % (pts, " ".join('%s="%s"' % (k.replace("_", "-"), a) for k, a in kw.items()))
        if cell is None:
            q05, q95 = series[1], series[3]
            ok = np.isfinite(q05) & np.isfinite(q95)
            band = " ".join("%.1f,%.1f" % (x, y) for x, y in zip(X(t[ok]), Y(q05[ok]))) + " " + \
                   " ".join("%.1f,%.1f" % (x, y) for x, y in zip(X(t[ok])[::-1], Y(q95[ok])[::-1]))
            out.append('<polygon class="band" points="%s" fill="#9ecae1" fill-opacity="0.6" stroke="none"/>' % band)
            out.append(poly(series[0], stroke="#9ecae1", stroke_width=1))
            out.append(poly(series[4], stroke="#9ecae1", stroke_width=1))
            out.append(poly(series[2], stroke="#08519c", stroke_width=2))
        else:
            out.append(poly(series[0], stroke="#0072bd", stroke_width=2))
        for r, lab in refs:
            black = key == "phise" and lab == ""
            out.append('<line class="limit" x1="%d" x2="%d" y1="%.1f" y2="%.1f" stroke="%s" stroke-width="1.5"%s/>' %
                       (L, L + w, Y(r), Y(r), "black" if black else "red", "" if black else ' stroke-dasharray="6 4"'))
            if lab:
                out.append('<text x="%d" y="%.1f" fill="red" text-anchor="end">%s</text>' % (L + w - 4, Y(r) - 3, lab))
        out.append("</g>")
    out.append('<text x="%d" y="%d" text-anchor="middle">Time [s]</text></svg>' % (W // 2, H - 6))
    svg = "\n".join(out)
    if path:
        with open(path, "w") as f:
            f.write(svg)
    return svg
