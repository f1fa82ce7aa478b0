"""Host model of the disk's transient structures (filaments, hotspots, RT spikes).

Reference: EntityInstance / EntityFactory (render.py:493-792), _spawn_single_* (1667-1866),
_init_lifecycle_system / _advance_lifecycle_frame (4079-4153).  The model is driven by three
numpy PCG64 streams; the ORDER of the draws is the contract and is kept:

  filament : source_phi, r_pos, sigma_r, sigma_phi0, peak_density, temp_ratio
  hotspot  : phi, r_rand, phi_width, r_width_extra, intensity_extra, (unused power(0.4) draw)
  rt_spike : phi, r_base_u, phi_width, r_length, intensity, delta_T
  then, for every spawn: lifetime, and the four fade-noise draws (freq1, freq2, phase1, phase2)

Unlike the reference, entities keep only their PARAMETERS: the per-row azimuthal profiles are
evaluated by the device kernel (csrc/texture.cu, entity_accumulate_kernel) instead of being
tabulated on the host and rolled with numpy every frame (0.71 s/frame at the fhd texture).  The
tabulated arrays remain available lazily (`phi_density`, `phi_temp`, `fade_noise`) for callers
that inspect them.
"""
import math
from typing import List, Tuple

import numpy as np

from . import _lib as L

FILAMENT_SHEAR_ALPHA = 0.1
FILAMENT_TAU_COOL = 50.0
FILAMENT_DEATH_THRESHOLD = 0.008
FILAMENT_MAX_LIFETIME = 120.0
FILAMENT_BIRTH_FADE_DUR = 5.0

KIND = {"filament": 0, "hotspot": 1, "rt_spike": 2}
TWO_PI = 2 * np.pi


def _rows_where(mask, r_norm_all, fallback_r):
    idx = np.flatnonzero(mask)
    if idx.size == 0:
        c = int(np.argmin(np.abs(r_norm_all - fallback_r)))
        return c, c + 1
    return int(idx[0]), int(idx[-1]) + 1


def _nearest_omega(r_norm_all, omega_all, r):
    return float(omega_all[int(np.argmin(np.abs(r_norm_all - r)))])


def draw_filament(rng, r_norm_all, omega_all):
    """Six draws; a circular Gaussian blob later sheared into an arc (render.py:1703-1722)."""
    source_phi = float(rng.uniform(0, TWO_PI))
    r_pos = float(rng.uniform(0.05, 0.95))
    base_r = 0.05 + r_pos ** 0.6 * 0.9
    sigma_r = float(rng.uniform(0.005, 0.015))
    sigma_phi0 = float(rng.uniform(0.04, 0.10))
    peak_density = float(rng.uniform(0.5, 1.0))
    peak_temp = peak_density * float(rng.uniform(0.15, 0.35))
    rows = _rows_where(np.abs(r_norm_all - base_r) < 4 * sigma_r, r_norm_all, base_r)
    return dict(rows=rows, omega=_nearest_omega(r_norm_all, omega_all, base_r), source_phi=source_phi,
                base_r=base_r, sigma_r=sigma_r, sigma_phi0=sigma_phi0, peak_density=peak_density,
                peak_temp=peak_temp)


def draw_hotspot(rng, r_norm_all, omega_all):
    """Six draws; von-Mises x Gaussian bright patch (render.py:1754-1793)."""
    phi0 = float(rng.uniform(0, TWO_PI))
    r_rand = float(rng.uniform(0, 1))
    h_r = 0.1 + r_rand ** 0.6 * 0.85
    phi_width = float(rng.uniform(0.08, 0.20))
    r_width = 0.02 + float(rng.uniform(0, 0.03))
    intensity = 0.3 + (1 - h_r) * 0.6 + float(rng.uniform(0, 0.1))
    rng.power(0.4)                                  # h_delta_T: drawn by the reference, never used
    lo, hi = h_r - 3 * r_width, h_r + 3 * r_width
    rows = _rows_where((r_norm_all >= lo) & (r_norm_all <= hi), r_norm_all, h_r)
    return dict(rows=rows, omega=_nearest_omega(r_norm_all, omega_all, h_r), phi0=phi0, r0=h_r,
                phi_width=phi_width, r_width=r_width, intensity=intensity)


def draw_rt_spike(rng, r_norm_all, omega_all):
    """Six draws; radial Rayleigh-Taylor finger near the inner edge (render.py:1825-1866)."""
    phi0 = float(rng.uniform(0, TWO_PI))
    r_base = float(np.power(rng.uniform(0.01, 0.15), 1.5))
    phi_width = float(rng.uniform(0.08, 0.20))
    r_length = float(rng.uniform(0.08, 0.20))
    intensity = float(rng.uniform(0.8, 1.0))
    delta_T = float(rng.uniform(0.5, 1.2))
    lo, hi = max(r_base - 0.02, 0.0), r_base + r_length * 2.5
    rows = _rows_where((r_norm_all >= lo) & (r_norm_all <= hi), r_norm_all, r_base)
    return dict(rows=rows, omega=_nearest_omega(r_norm_all, omega_all, r_base + r_length * 0.5),
                phi0=phi0, r0=r_base, phi_width=phi_width, r_length=r_length, intensity=intensity,
                delta_T=delta_T)


_DRAW = {"filament": draw_filament, "hotspot": draw_hotspot, "rt_spike": draw_rt_spike}


def _profiles(kind, q, r_norm_all, n_phi):
    """Tabulate (phi_density, phi_temp) like the reference's spawn functions (float32 rows)."""
    r0, r1 = q["rows"]
    if kind == "filament":
        e = np.empty((0, 0), dtype=np.float32)
        return e, e
    phi = np.linspace(0, TWO_PI, n_phi, endpoint=False)
    kappa = 1.5 / (q["phi_width"] ** 2)
    phi_prof = np.exp(kappa * (np.cos(phi - q["phi0"]) - 1))
    dens = np.zeros((r1 - r0, n_phi), dtype=np.float32)
    temp = np.zeros((r1 - r0, n_phi), dtype=np.float32)
    for k, ri in enumerate(range(r0, r1)):
        d = r_norm_all[ri] - q["r0"]
        if kind == "hotspot":
            radial = np.exp(-0.5 * (d / (q["r_width"] + 1e-8)) ** 2)
            dens[k] = phi_prof * radial * q["intensity"]
            temp[k] = dens[k] * 0.12
        else:
            fade_out = np.clip(q["r_length"] * 2 - d, 0, 1)
            fade_in = np.clip(d / (q["r_length"] * 0.3 + 1e-8), 0, 1)
            radial = np.exp(-0.5 * (d / (q["r_length"] * 0.4 + 1e-8)) ** 2) * fade_out * fade_in
            dens[k] = phi_prof * radial * q["intensity"]
            temp[k] = dens[k] * q["delta_T"]
    dens = np.clip(dens, 0, 1)
    if kind == "hotspot":
        temp = np.clip(temp, 0, 1)
    return dens, temp


class EntityInstance:
    """One transient structure; same attribute / method surface as the reference's dataclass."""

    def __init__(self, entity_type, q, birth_time, lifetime, fade_in, fade_out, fade_spec,
                 r_norm_all, n_phi):
        self.entity_type = entity_type
        self.q = q
        self.birth_time = birth_time
        self.lifetime = lifetime
        self.fade_in = fade_in
        self.fade_out = fade_out
        self.omega = q["omega"]
        self._fade_spec = fade_spec
        self._r_norm_all = r_norm_all
        self._n_phi = n_phi
        self._tab = None
        self.row_indices = np.arange(q["rows"][0], q["rows"][1])
        f = entity_type == "filament"
        self.source_phi = q["source_phi"] if f else 0.0
        self.total_extent = TWO_PI if f else 0.0
        self.alpha_shear = FILAMENT_SHEAR_ALPHA * self.omega if f else 0.0
        self.tau_cool = FILAMENT_TAU_COOL
        self.blob_base_r = q["base_r"] if f else 0.0
        self.blob_sigma_r = q["sigma_r"] if f else 0.0
        self.blob_sigma_phi0 = q["sigma_phi0"] if f else 0.0
        self.blob_peak_density = q["peak_density"] if f else 0.0
        self.blob_peak_temp = q["peak_temp"] if f else 0.0

    # lazily tabulated arrays (never needed by the device path)
    def _tabulate(self):
        if self._tab is None:
            self._tab = _profiles(self.entity_type, self.q, self._r_norm_all, self._n_phi)
        return self._tab

    @property
    def phi_density(self):
        return self._tabulate()[0]

    @property
    def phi_temp(self):
        return self._tabulate()[1]

    @property
    def fade_noise(self):
        f1, f2, p1, p2 = self._fade_spec
        phi = np.linspace(0, TWO_PI, self._n_phi, endpoint=False)
        n = 0.6 * np.sin(phi * f1 + p1) + 0.4 * np.sin(phi * f2 + p2)
        return np.clip(n * 0.5 + 0.5, 0, 1).astype(np.float32)

    @property
    def total_duration(self):
        return self.fade_in + self.lifetime + self.fade_out

    def density_factor(self, age):
        """Filament shear dilution x radiative cooling: (s0 / (s0 + a*age)) * exp(-age / tau)."""
        s0 = max(self.blob_sigma_phi0, 1e-6)
        shear = s0 / (s0 + self.alpha_shear * age)
        cool = math.exp(-age / self.tau_cool) if self.tau_cool > 0 else 1.0
        return shear * cool

    def is_dead(self, now):
        age = now - self.birth_time
        if self.entity_type == "filament":
            if age >= FILAMENT_MAX_LIFETIME:
                return True
            return age >= 0 and self.density_factor(age) < FILAMENT_DEATH_THRESHOLD
        return age >= self.total_duration

    def fade_factor(self, now):
        """Linear fade-in / plateau / fade-out envelope of hotspots and RT spikes."""
        age = now - self.birth_time
        if age < 0:
            return 0.0
        if age < self.fade_in:
            return age / self.fade_in if self.fade_in > 0 else 1.0
        rest = age - self.fade_in
        if rest < self.lifetime:
            return 1.0
        rest -= self.lifetime
        if rest < self.fade_out:
            return 1.0 - rest / self.fade_out if self.fade_out > 0 else 0.0
        return 0.0


class EntityFactory:
    """Keeps `target_count` entities alive: culls the dead, spawns at target/avg_lifetime per
    second (render.py:624-792).  `spawn_fn` is accepted for signature compatibility; the draw
    routine is selected by `entity_type`."""

    def __init__(self, spawn_fn, target_count: int, lifetime_range: Tuple[float, float],
                 fade_in: float, fade_out: float, n_r: int, n_phi: int, r_norm_all, omega_all,
                 seed: int = 0, entity_type: str = "generic"):
        self.spawn_fn = spawn_fn
        self.target_count = target_count
        self.lifetime_range = lifetime_range
        self.fade_in = fade_in
        self.fade_out = fade_out
        self.n_r = n_r
        self.n_phi = n_phi
        self.r_norm_all = r_norm_all
        self.omega_all = omega_all
        self.rng = np.random.default_rng(seed)
        self.entities: List[EntityInstance] = []
        self._spawn_debt = 0.0
        self._version = 0          # bumped whenever the population changes (SoA cache key)
        self._soa = None
        self._cull = None          # column cache of the filament cull: [list identity, birth, s0, alpha, tau]
        self._sm = None            # static-column matrix of the packer, kept in step with the list: [list identity, m, birth]
        self.entity_type = entity_type
        if entity_type not in _DRAW:
            raise ValueError(f"unknown entity_type {entity_type!r}")

    def _spawn_one(self, now):
        q = _DRAW[self.entity_type](self.rng, self.r_norm_all, self.omega_all)
        lifetime = float(self.rng.uniform(*self.lifetime_range))
        fade_spec = (int(self.rng.integers(3, 8)), int(self.rng.integers(8, 16)),
                     float(self.rng.uniform(0, TWO_PI)), float(self.rng.uniform(0, TWO_PI)))
        return EntityInstance(self.entity_type, q, now, lifetime, self.fade_in, self.fade_out,
                              fade_spec, self.r_norm_all, self.n_phi)

    @staticmethod
    def _filament_death_age(entity):
        for t in range(1, int(FILAMENT_MAX_LIFETIME) + 1):
            if entity.density_factor(float(t)) < FILAMENT_DEATH_THRESHOLD:
                return float(t)
        return FILAMENT_MAX_LIFETIME

    def seed_initial(self, now):
        """Start at steady state: spawn target_count entities with staggered ages."""
        n = max(self.target_count, 1)
        for i in range(self.target_count):
            e = self._spawn_one(now)
            if e.entity_type == "filament":
                span = max(self._filament_death_age(e) - FILAMENT_BIRTH_FADE_DUR, 1.0)
                stagger = FILAMENT_BIRTH_FADE_DUR + span * (i / n)
            else:
                stagger = (e.fade_in + e.lifetime) * (i / n)
            e.birth_time = now - stagger
            self.entities.append(e)
        self._version += 1

    def tick(self, now, dt):
        before = len(self.entities)
        if self.entity_type == "filament":
            self._cull_filaments(now)
        else:
            keep = [now - e.birth_time < e.fade_in + e.lifetime + e.fade_out for e in self.entities]
            if not all(keep):
                old = self.entities
                self.entities = [e for e, k in zip(old, keep) if k]
                self._sm_cull(old, np.array(keep, dtype=bool))
        if len(self.entities) != before:
            self._version += 1
        deficit = self.target_count - len(self.entities)
        if deficit <= 0:
            return
        rate = self.target_count / (sum(self.lifetime_range) / 2.0)
        self._spawn_debt += rate * dt
        n_spawn = min(int(self._spawn_debt), deficit)
        self._spawn_debt -= n_spawn
        for _ in range(n_spawn):
            e = self._spawn_one(now)
            self.entities.append(e)
            self._sm_append(e)
            if self.entity_type == "filament":
                self._cull_append(e)
        if n_spawn:
            self._version += 1

    def _cull_filaments(self, now):
        """EntityInstance.is_dead for every filament (render.py:560-621 semantics).  The decision
        `shear * cool < threshold` is evaluated vectorised (numpy exp) on a column cache that is
        kept in step with the list; the few entities within 1e-12 of the threshold -- where
        numpy's exp and math.exp could disagree in the last ulp -- are re-decided with the scalar
        arithmetic the reference uses, so the population (and with it the RNG call order) is
        exactly the reference's."""
        ents = self.entities
        c = self._cull
        if c is None or c[0] is not ents or len(c[1]) != len(ents):
            n = len(ents)
            tau = np.array([e.tau_cool for e in ents], dtype=np.float64).reshape(n)
            c = self._cull = [ents,
                              np.array([e.birth_time for e in ents], dtype=np.float64).reshape(n),
                              np.array([max(e.blob_sigma_phi0, 1e-6) for e in ents], dtype=np.float64).reshape(n),
                              np.array([e.alpha_shear for e in ents], dtype=np.float64).reshape(n),
                              np.where(tau > 0, -1.0 / np.where(tau > 0, tau, 1.0), 0.0)]     # -1 / tau (0: no cooling)
        if not ents:
            return
        _, birth, s0, alpha, neg_inv_tau = c
        age = now - birth
        val = s0 / (s0 + alpha * age) * np.exp(age * neg_inv_tau)
        thr, tmax = FILAMENT_DEATH_THRESHOLD, FILAMENT_MAX_LIFETIME
        live = age >= 0
        dead = (age >= tmax) | (live & (val < thr))
        close = live & (np.abs(val - thr) < 1e-12)
        for i in (np.nonzero(close)[0] if close.any() else ()):
            e, a = ents[i], now - ents[i].birth_time
            if a >= tmax:
                continue
            cl = math.exp(-a / e.tau_cool) if e.tau_cool > 0 else 1.0
            dead[i] = (max(e.blob_sigma_phi0, 1e-6) / (max(e.blob_sigma_phi0, 1e-6) + e.alpha_shear * a)) * cl < thr
        if dead.any():
            keep = ~dead
            self.entities = [e for e, k in zip(ents, keep) if k]
            self._cull = [self.entities, birth[keep], s0[keep], alpha[keep], neg_inv_tau[keep]]
            self._sm_cull(ents, keep)

    # the packer's static columns (lifecycle._factory_soa), updated in place of a rebuild from the Python objects
    def _sm_cull(self, old_list, keep):
        c = self._sm
        if c is not None and c[0] is old_list and len(c[1]) == len(old_list):
            self._sm = [self.entities, c[1][keep], c[2][keep]]
        else:
            self._sm = None

    def _sm_append(self, e):
        c = self._sm
        if c is not None and c[0] is self.entities and len(c[1]) == len(self.entities) - 1:
            c[1] = np.concatenate([c[1], np.array([_static_row(e)], dtype=np.float64)])
            c[2] = np.append(c[2], e.birth_time)
        else:
            self._sm = None

    def _cull_append(self, e):
        c = self._cull
        if c is not None and c[0] is self.entities and len(c[1]) == len(self.entities) - 1:
            c[1] = np.append(c[1], e.birth_time); c[2] = np.append(c[2], max(e.blob_sigma_phi0, 1e-6))
            c[3] = np.append(c[3], e.alpha_shear)
            c[4] = np.append(c[4], -1.0 / e.tau_cool if e.tau_cool > 0 else 0.0)

    @property
    def alive_entities(self):
        return self.entities


# API-compatible spawn functions (reference: _spawn_single_*, render.py:1667-1866)
def _spawn_single_filament(rng, n_r, n_phi, r_norm_all, omega_all):
    q = draw_filament(rng, r_norm_all, omega_all)
    e = np.empty((0, 0), dtype=np.float32)
    return (np.arange(*q["rows"]), e, e, q["omega"], q["source_phi"], TWO_PI, q["sigma_r"],
            q["sigma_phi0"], q["peak_density"], q["peak_temp"], q["base_r"])


def _spawn_single_hotspot(rng, n_r, n_phi, r_norm_all, omega_all):
    q = draw_hotspot(rng, r_norm_all, omega_all)
    d, t = _profiles("hotspot", q, r_norm_all, n_phi)
    return np.arange(*q["rows"]), d, t, q["omega"]


def _spawn_single_rt_spike(rng, n_r, n_phi, r_norm_all, omega_all):
    q = draw_rt_spike(rng, r_norm_all, omega_all)
    d, t = _profiles("rt_spike", q, r_norm_all, n_phi)
    return np.arange(*q["rows"]), d, t, q["omega"]


def pack_entities(factories, now, n_r):
    """Per-frame scalars of accumulate_entity_layer (render.py:3605-3642), one bhr_entity each.
    All host arithmetic is Python float64 exactly as in the reference; per-texel work is the
    device kernel's."""
    out = []
    for key in ("filament", "rt_spike", "hotspot"):
        f = factories.get(key)
        if f is None:
            continue
        for e in f.alive_entities:
            age = now - e.birth_time
            ent = L.BhrEntity()
            ent.kind = KIND[e.entity_type]
            ent.row_begin = max(int(e.q["rows"][0]), 0)
            ent.row_end = min(int(e.q["rows"][1]), n_r)
            ent.age = age
            if e.entity_type == "filament":
                if e.density_factor(age) < FILAMENT_DEATH_THRESHOLD:
                    continue
                s0 = max(e.blob_sigma_phi0, 1e-6)
                sigma_t = s0 + e.alpha_shear * age
                amp_d = e.blob_peak_density * s0 / sigma_t
                amp_t = e.blob_peak_temp * s0 / sigma_t
                birth = min(age / FILAMENT_BIRTH_FADE_DUR, 1.0) if FILAMENT_BIRTH_FADE_DUR > 0 else 1.0
                cool = math.exp(-age / e.tau_cool) if e.tau_cool > 0 else 1.0
                sigma_r = max(e.blob_sigma_r, 1e-6)
                ent.scale = birth * cool
                vals = [e.source_phi, e.blob_base_r, 0.5 / (sigma_r * sigma_r),
                        0.5 / (sigma_t * sigma_t), amp_d * birth * cool, amp_t * birth * cool]
            else:
                alpha = e.fade_factor(now)
                if alpha <= 0:
                    continue
                ent.scale = alpha
                q = e.q
                if e.entity_type == "hotspot":
                    vals = [q["phi0"], q["r0"], q["phi_width"], q["r_width"], q["intensity"]]
                else:
                    vals = [q["phi0"], q["r0"], q["phi_width"], q["r_length"], q["intensity"],
                            q["delta_T"]]
            for i, v in enumerate(vals):
                ent.p[i] = v
            out.append(ent)
    return out


ENTITY_DTYPE = np.dtype([("kind", "<i4"), ("row_begin", "<i4"), ("row_end", "<i4"), ("_pad", "<i4"),
                         ("age", "<f8"), ("scale", "<f8"), ("p", "<f8", (8,))])
assert ENTITY_DTYPE.itemsize == 96


_FIL_COLS = ("source_phi", "alpha_shear", "tau_cool", "blob_base_r", "blob_sigma_r", "blob_sigma_phi0",
             "blob_peak_density", "blob_peak_temp")
_OTHER_PARAMS = {"hotspot": ("phi0", "r0", "phi_width", "r_width", "intensity"),
                 "rt_spike": ("phi0", "r0", "phi_width", "r_length", "intensity", "delta_T")}


def _static_row(e):
    """The columns of an entity that never change after its birth (cached on the instance): rows, and the
    filament's blob parameters or the hotspot's / spike's envelope + profile parameters."""
    row = e.__dict__.get("_static")
    if row is None:
        if e.entity_type == "filament":
            row = (float(e.q["rows"][0]), float(e.q["rows"][1])) + tuple(float(getattr(e, k)) for k in _FIL_COLS)
        else:
            row = (float(e.q["rows"][0]), float(e.q["rows"][1]), float(e.fade_in), float(e.lifetime), float(e.fade_out)) + \
                  tuple(float(e.q[k]) for k in _OTHER_PARAMS[e.entity_type])
        e._static = row
    return row


def _factory_soa(f, n_r):
    """Static per-entity columns of a factory + the packed-entity template with every static field filled in,
    rebuilt only when its population changes (one C-level conversion of the cached per-entity rows).  The template
    is the bhr_entity array seen as 12 float64 slots per entity: slot 0 = (kind, row_begin) and slot 1 =
    (row_end, pad) as int32 pairs, slot 2 = age, slot 3 = scale, slots 4..11 = p[0..7]."""
    key = (f._version, len(f.entities), n_r)
    if f._soa is not None and f._soa[0] == key:
        return f._soa[1]
    ents = f.entities
    n = len(ents)
    sm = getattr(f, "_sm", None)
    if sm is not None and sm[0] is ents and len(sm[1]) == n and len(sm[2]) == n:
        m, birth = sm[1], sm[2]               # kept in step by the factory's tick (no pass over the Python objects)
    else:
        m = np.array([_static_row(e) for e in ents], dtype=np.float64).reshape(n, -1)
        birth = np.fromiter((e.birth_time for e in ents), dtype=np.float64, count=n)
        f._sm = [ents, m, birth]
    tmpl = np.zeros((n, 12), dtype=np.float64)
    ti = tmpl.view(np.int32)                     # (n, 24)
    ti[:, 1] = np.maximum(m[:, 0].astype(np.int32), 0)
    ti[:, 2] = np.minimum(m[:, 1].astype(np.int32), n_r)
    col = dict(birth=birth, tmpl=tmpl)
    if f.entity_type == "filament":
        for k, name in enumerate(_FIL_COLS):
            col[name] = np.ascontiguousarray(m[:, 2 + k])
        sigma_r = np.maximum(col["blob_sigma_r"], 1e-6)
        col["s0"] = np.maximum(col["blob_sigma_phi0"], 1e-6)
        col["tau_safe"] = np.where(col["tau_cool"] > 0, col["tau_cool"], 1.0)
        col["no_tau"] = None if (col["tau_cool"] > 0).all() else ~(col["tau_cool"] > 0)
        ti[:, 0] = 0
        tmpl[:, 4] = col["source_phi"]
        tmpl[:, 5] = col["blob_base_r"]
        tmpl[:, 6] = 0.5 / (sigma_r * sigma_r)
    else:
        col["fade_in"], col["life"], col["fade_out"] = (np.ascontiguousarray(m[:, 2 + k]) for k in range(3))
        ti[:, 0] = KIND[f.entity_type]
        tmpl[:, 4:4 + m.shape[1] - 5] = m[:, 5:]
    f._soa = (key, col)
    return col


def pack_entities_array(factories, now, n_r):
    """Vectorised `pack_entities`: a numpy array with bhr_entity's layout.  Same formulas in
    float64 (numpy's exp instead of math.exp: <= 1 ulp of a double on the scale factors)."""
    parts = []
    for key in ("filament", "rt_spike", "hotspot"):
        f = factories.get(key)
        if f is None or not f.entities:
            continue
        c = _factory_soa(f, n_r)
        age = now - c["birth"]
        out = c["tmpl"].copy()
        if f.entity_type == "filament":
            s0 = c["s0"]
            sigma_t = s0 + c["alpha_shear"] * age
            cool = np.exp(-age / c["tau_safe"])
            if c["no_tau"] is not None:
                cool[c["no_tau"]] = 1.0
            shear = s0 / sigma_t
            alive = shear * cool >= FILAMENT_DEATH_THRESHOLD
            birth = np.minimum(age / FILAMENT_BIRTH_FADE_DUR, 1.0)
            out[:, 3] = birth * cool
            out[:, 7] = 0.5 / (sigma_t * sigma_t)
            out[:, 8] = c["blob_peak_density"] * s0 / sigma_t * birth * cool
            out[:, 9] = c["blob_peak_temp"] * s0 / sigma_t * birth * cool
        else:
            fin, life, fout = c["fade_in"], c["life"], c["fade_out"]
            rest = age - fin
            with np.errstate(divide="ignore", invalid="ignore"):
                ramp_in = np.where(fin > 0, age / np.where(fin > 0, fin, 1.0), 1.0)
                ramp_out = np.where(fout > 0, 1.0 - (rest - life) / np.where(fout > 0, fout, 1.0), 0.0)
            alpha = np.where(age < 0, 0.0, np.where(age < fin, ramp_in, np.where(
                rest < life, 1.0, np.where(rest - life < fout, ramp_out, 0.0))))
            alive = alpha > 0
            out[:, 3] = alpha
        out[:, 2] = age
        parts.append(out if alive.all() else out[alive])
    if not parts:
        return np.zeros(0, dtype=ENTITY_DTYPE)
    return np.ascontiguousarray(np.concatenate(parts)).view(ENTITY_DTYPE).reshape(-1)


def _row_runs(row_indices, n_r):
    """Contiguous [begin, end) runs of an entity's row_indices clipped to the texture, with the index of each run's
    first row in the entity's profile arrays."""
    rows = np.asarray(row_indices, dtype=np.int64)
    runs, k = [], 0
    while k < len(rows):
        j = k
        while j + 1 < len(rows) and rows[j + 1] == rows[j] + 1:
            j += 1
        a, b = int(rows[k]), int(rows[j]) + 1
        lo, hi = max(a, 0), min(b, n_r)
        if lo < hi:
            runs.append((lo, hi, k + (lo - a)))
        k = j + 1
    return runs


def pack_foreign_entities(factories, now, n_r, n_phi, table_cache):
    """bhr_entity array for factories whose entities are NOT this module's EntityInstance -- e.g. the reference's
    own EntityFactory / EntityInstance objects (render.py:499-792), which carry tabulated (rows, n_phi) float32
    profiles instead of analytic parameters.  Filaments are analytic in the reference too (render.py:3605-3636)
    and are packed from their attributes; hotspots and RT spikes become kind 3 / 4 entries that reference their
    phi_density / phi_temp rows in one float32 table buffer.  `table_cache`: dict kept by the renderer; returns
    (entities, tables or None): tables is the new buffer to upload when the set of tabulated entities changed."""
    ents = []
    tabulated = []
    for key in ("filament", "rt_spike", "hotspot"):
        f = factories.get(key)
        if f is None:
            continue
        for e in f.alive_entities:
            age = now - e.birth_time
            if e.entity_type == "filament":
                if e.density_factor(age) < FILAMENT_DEATH_THRESHOLD:
                    continue
                s0 = max(e.blob_sigma_phi0, 1e-6)
                sigma_t = s0 + e.alpha_shear * age
                birth = min(age / FILAMENT_BIRTH_FADE_DUR, 1.0) if FILAMENT_BIRTH_FADE_DUR > 0 else 1.0
                cool = math.exp(-age / e.tau_cool) if e.tau_cool > 0 else 1.0
                sigma_r = max(e.blob_sigma_r, 1e-6)
                vals = [e.source_phi, e.blob_base_r, 0.5 / (sigma_r * sigma_r), 0.5 / (sigma_t * sigma_t),
                        e.blob_peak_density * s0 / sigma_t * birth * cool, e.blob_peak_temp * s0 / sigma_t * birth * cool]
                for lo, hi, _ in _row_runs(e.row_indices, n_r):
                    ents.append((0, lo, hi, age, birth * cool, vals))
            else:
                alpha = e.fade_factor(now)
                if alpha <= 0:
                    continue
                tabulated.append((e, 3 if key == "hotspot" else 4, age, alpha))
    # table buffer: rebuilt (and re-uploaded) only when the set of tabulated entities changes.  Every contiguous run
    # of an entity's rows inside the texture gets its density rows followed by its temperature rows.
    ids = tuple(id(e) for e, _, _, _ in tabulated)
    tables = None
    if table_cache.get("ids") != ids:
        chunks, runs_of, off = [], {}, 0
        for e, _, _, _ in tabulated:
            d = np.ascontiguousarray(e.phi_density, dtype=np.float32).reshape(-1, n_phi)
            t = np.ascontiguousarray(e.phi_temp, dtype=np.float32).reshape(-1, n_phi)
            mine = []
            for lo, hi, k0 in _row_runs(e.row_indices, n_r):
                mine.append((lo, hi, off))
                chunks += [d[k0:k0 + hi - lo].ravel(), t[k0:k0 + hi - lo].ravel()]
                off += 2 * (hi - lo) * n_phi
            runs_of[id(e)] = mine
        tables = np.concatenate(chunks) if chunks else np.zeros(0, dtype=np.float32)
        table_cache.update(ids=ids, runs=runs_of, keep=[e for e, _, _, _ in tabulated])   # (keep: ids stay unique while cached)
    for e, kind, age, alpha in tabulated:
        for lo, hi, off in table_cache["runs"][id(e)]:
            ents.append((kind, lo, hi, age, alpha, [float(off)]))
    out = np.zeros(len(ents), dtype=ENTITY_DTYPE)
    for i, (kind, lo, hi, age, scale, vals) in enumerate(ents):
        out[i]["kind"], out[i]["row_begin"], out[i]["row_end"], out[i]["age"], out[i]["scale"] = kind, lo, hi, age, scale
        out[i]["p"][:len(vals)] = vals
    return out, tables


def init_lifecycle_system(renderer, n_r, n_phi, seed=42):
    """_init_lifecycle_system (render.py:4079-4130): background parameters, three factories
    (seeds +100/+200/+300, targets 200/30/15), staggered initial population, first texture."""
    renderer.init_background_layer(n_r=n_r, n_phi=n_phi, seed=seed)
    factories = make_factories(renderer.r_disk_inner, renderer.r_disk_outer, n_r, n_phi, seed)
    renderer.generate_background(t=0.0)
    renderer.accumulate_entity_layer(factories, now=0.0)
    renderer.recompute_interactive_stats()
    renderer.compose_interactive_texture()
    return factories


def make_factories(r_disk_inner, r_disk_outer, n_r, n_phi, seed=42):
    """The three seeded, pre-populated factories of _init_lifecycle_system (host only)."""
    r_norm_all = np.linspace(0, 1, n_r)
    r_vals = r_disk_inner + (r_disk_outer - r_disk_inner) * r_norm_all
    omega_all = np.sqrt(0.5 / (r_vals ** 3 + 1e-6)).astype(np.float32)
    spec = [("filament", _spawn_single_filament, 200, (15.0, 60.0), 0.0, 0.0, 100),
            ("hotspot", _spawn_single_hotspot, 30, (15.0, 30.0), 4.0, 4.0, 200),
            ("rt_spike", _spawn_single_rt_spike, 15, (15.0, 30.0), 3.0, 3.0, 300)]
    factories = {}
    for name, fn, target, life, fin, fout, ds in spec:
        factories[name] = EntityFactory(fn, target_count=target, lifetime_range=life, fade_in=fin,
                                        fade_out=fout, n_r=n_r, n_phi=n_phi, r_norm_all=r_norm_all,
                                        omega_all=omega_all, seed=seed + ds, entity_type=name)
    for f in factories.values():
        f.seed_initial(now=0.0)
    return factories


def advance_lifecycle_frame(renderer, factories, t, dt, recompute_stats=False, solo_idx=-1):
    """_advance_lifecycle_frame (render.py:4133-4153): tick, background, entities, [stats], compose."""
    for f in factories.values():
        f.tick(now=t, dt=dt)
    renderer.generate_background(t=t)
    renderer.accumulate_entity_layer(factories, now=t)
    if recompute_stats:
        renderer.recompute_interactive_stats()
    renderer.compose_interactive_texture(solo_idx=solo_idx)


_init_lifecycle_system = init_lifecycle_system
_advance_lifecycle_frame = advance_lifecycle_frame
