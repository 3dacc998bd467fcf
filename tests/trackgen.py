"""TEST INFRASTRUCTURE: adversarial track geometry for the parity tests (tests/test_gpu_adversarial.py).

The bench generator (csrc/synth_tracks.cpp) only makes smooth star-shaped curves with rings at a constant +-1.75 m.
The corridor code of the kernels (anchors, clearances, crossing parity, existence certificates, update path) rests on
geometric reasoning that such tracks barely stress, so these generators make the cases the reference's ray casting
(main.cpp:478-512, 694-711) treats without any assumption:

  * cone jitter        ring vertices displaced by Gaussian noise (sigma 5-20 cm): rings are no offset curves any more
  * variable width     half-width varying between 1.25 and 3 m along the track
  * hairpins           180-degree turns of radius 4-6 m: a normal ray that misses / passes its own boundary runs into the
                       far side of the same ring
  * near sections      two sections of the track whose boundaries are ~3 m apart
  * segment soup       rings given as unordered, randomly oriented segments (no vertex chain), mixed with chained rings

Everything is a pure function of its seed.  The checker is the oracle (oracle/raceline_oracle.c, pinned bit for bit
to the reference); the generator only has to produce the SAME bits for both sides, not a drivable track.
"""
from __future__ import annotations

import numpy as np


def _resample_closed(P, n):
    """n points uniformly spaced by arc length on the closed polyline P; returns (points, length)."""
    Q = np.vstack([P, P[:1]])
    d = np.hypot(*np.diff(Q, axis=0).T)
    s = np.concatenate([[0.0], np.cumsum(d)])
    L = s[-1]
    t = np.arange(n) * (L / n)
    return np.stack([np.interp(t, s, Q[:, 0]), np.interp(t, s, Q[:, 1])], axis=1), float(L)


def _normals(P):
    t = np.roll(P, -1, axis=0) - np.roll(P, 1, axis=0)
    t /= np.maximum(1e-12, np.hypot(t[:, 0], t[:, 1]))[:, None]
    return np.stack([-t[:, 1], t[:, 0]], axis=1)


def _arc(c, r, a0, a1, step=0.05):
    k = max(8, int(abs(a1 - a0) * r / step))
    a = np.linspace(a0, a1, k, endpoint=False)
    return np.stack([c[0] + r * np.cos(a), c[1] + r * np.sin(a)], axis=1)


def _line(p, q, step=0.05):
    k = max(2, int(np.hypot(q[0] - p[0], q[1] - p[1]) / step))
    t = np.linspace(0.0, 1.0, k, endpoint=False)[:, None]
    return np.asarray(p)[None, :] * (1 - t) + np.asarray(q)[None, :] * t


def centre_flower(rng, length):
    """Smooth star-shaped curve (the bench generator's family) of roughly the given length."""
    th = np.linspace(0, 2 * np.pi, 8192, endpoint=False)
    r = 1.0 + sum(rng.uniform(0.02, 0.08) * np.sin(k * th + rng.uniform(0, 6.28)) for k in rng.choice(np.arange(2, 9), 3, replace=False))
    P = np.stack([r * np.cos(th), r * np.sin(th)], axis=1)
    per = np.hypot(*np.diff(np.vstack([P, P[:1]]), axis=0).T).sum()
    return P * (length / per)


def centre_paperclip(rng, straight, radius):
    """Two straights joined by two hairpins of the given radius (counter-clockwise)."""
    r, s = radius, straight
    return np.vstack([_line((0, -r), (s, -r)), _arc((s, 0), r, -np.pi / 2, np.pi / 2), _line((s, r), (0, r)),
                      _arc((0, 0), r, np.pi / 2, 3 * np.pi / 2)])


def centre_serpentine(rng, n_turns, straight, radius):
    """A serpentine of hairpins closed by a long return straight: consecutive straights are 2*radius apart."""
    r, s = radius, straight
    pts, y = [], 0.0
    for k in range(n_turns):
        if k % 2 == 0:
            pts += [_line((0, y), (s, y)), _arc((s, y + r), r, -np.pi / 2, np.pi / 2)]
        else:
            pts += [_line((s, y), (0, y)), _arc((0, y + r), r, -np.pi / 2, -3 * np.pi / 2)]
        y += 2 * r
    # close: from the end of the last hairpin around the outside back to the start
    R = 2.5 * r
    if n_turns % 2 == 0:     # we are at (0, y) heading +x ... go left around
        pts += [_line((0, y), (s, y)), _arc((s, y - R), R, np.pi / 2, -np.pi / 2), _line((s, y - 2 * R), (s + 0.0, y - 2 * R))]
        x_end, y_end = s, y - 2 * R
        pts += [_line((x_end, y_end), (s, -R * 0 - 0.0))] if False else []
        P = np.vstack(pts)
        # generic closing: straight line back to the first point through two wide arcs is overkill; use a big detour
        a, b = P[-1], P[0]
        detour = np.array([[a[0] + 4 * R, a[1]], [a[0] + 4 * R, b[1] - 4 * R], [b[0] - 4 * R, b[1] - 4 * R], [b[0] - 4 * R, b[1]]])
        return np.vstack([P, _smooth_path(np.vstack([a, detour, b]))])
    P = np.vstack(pts)
    a, b = P[-1], P[0]
    detour = np.array([[a[0] - 4 * R, a[1]], [a[0] - 4 * R, b[1] - 4 * R], [b[0] + s + 4 * R, b[1] - 4 * R], [b[0] + s + 4 * R, b[1] - 2 * R],
                       [b[0] - 2 * R, b[1] - 2 * R], [b[0] - 2 * R, b[1]]])
    return np.vstack([P, _smooth_path(np.vstack([a, detour, b]))])


def _smooth_path(ctrl, step=0.05, rounds=6):
    """Polyline through control points with Chaikin corner cutting (keeps the end points); dense resample."""
    P = np.asarray(ctrl, dtype=float)
    for _ in range(rounds):
        Q = [P[0]]
        for a, b in zip(P[:-1], P[1:]):
            Q += [0.75 * a + 0.25 * b, 0.25 * a + 0.75 * b]
        Q.append(P[-1])
        P = np.array(Q)
    d = np.hypot(*np.diff(P, axis=0).T)
    s = np.concatenate([[0], np.cumsum(d)])
    t = np.arange(0, s[-1], step)
    return np.stack([np.interp(t, s, P[:, 0]), np.interp(t, s, P[:, 1])], axis=1)[1:]


def centre_hourglass(rng, lobe, gap):
    """Two lobes joined by a waist whose two sections run `gap` metres apart (centre to centre)."""
    g = gap / 2
    ctrl = np.array([[-lobe, -lobe], [-2, -g], [2, -g], [lobe, -lobe], [1.6 * lobe, 0], [lobe, lobe], [2, g], [-2, g], [-lobe, lobe],
                     [-1.6 * lobe, 0], [-lobe, -lobe]], dtype=float)
    P = ctrl
    for _ in range(7):   # closed Chaikin
        Q = []
        for a, b in zip(P, np.roll(P, -1, axis=0)):
            Q += [0.75 * a + 0.25 * b, 0.25 * a + 0.75 * b]
        P = np.array(Q)
    return P


def make_track(seed, n, kind="flower", m=None, jitter=0.0, width=(1.75, 1.75), soup=(False, False), reverse_outer=False):
    """Returns (center (n,2), inner_seg (M,4), outer_seg (M,4), L).

    kind: flower | paperclip | serpentine | hourglass; jitter: sigma of the cone noise in metres; width: (min, max)
    half-width in metres, varying smoothly along the track; soup: per ring, emit the segments shuffled and randomly
    flipped instead of as a vertex chain."""
    rng = np.random.default_rng(seed)
    h = rng.uniform(1.5, 1.9)
    length = n * h
    if kind == "flower":
        raw = centre_flower(rng, length)
    elif kind == "paperclip":
        r = rng.uniform(4.0, 6.0)
        raw = centre_paperclip(rng, max(4.0, (length - 2 * np.pi * r) / 2), r)
    elif kind == "serpentine":
        r = rng.uniform(4.0, 6.0)
        turns = int(rng.integers(3, 6))
        raw = centre_serpentine(rng, turns, max(8.0, length / (2.2 * turns + 6)), r)
    elif kind == "hourglass":
        raw = centre_hourglass(rng, max(12.0, length / 9.0), gap=2 * width[1] + 3.0)
    else:
        raise ValueError(kind)
    center, L = _resample_closed(raw, n)
    m = int(round(n / 2.2)) if m is None else int(m)
    cones, _ = _resample_closed(raw, m)
    nrm = _normals(cones)
    s = np.arange(m) / m
    w = width[0] + (width[1] - width[0]) * 0.5 * (1 + np.sin(2 * np.pi * (3 * s + rng.uniform())) * np.cos(2 * np.pi * (s + rng.uniform())))
    inner = cones + nrm * w[:, None] + rng.normal(0.0, jitter, (m, 2)) if jitter > 0 else cones + nrm * w[:, None]
    w2 = width[0] + (width[1] - width[0]) * 0.5 * (1 + np.sin(2 * np.pi * (2 * s + rng.uniform())))
    outer = cones - nrm * w2[:, None] + (rng.normal(0.0, jitter, (m, 2)) if jitter > 0 else 0.0)

    def ring(V, as_soup, rev):
        if rev:
            V = V[::-1]
        E = np.concatenate([V, np.roll(V, -1, axis=0)], axis=1)
        if as_soup:
            E = E[rng.permutation(len(E))]
            flip = rng.random(len(E)) < 0.5
            E[flip] = E[flip][:, [2, 3, 0, 1]]
        return np.ascontiguousarray(E)

    return np.ascontiguousarray(center), ring(inner, soup[0], False), ring(outer, soup[1], reverse_outer), L
