"""Procedural equirectangular star field (host input of the render path).

Reference: generate_skybox / load_or_generate_skybox / _blackbody_rgb (render.py:136-368).  The
picture is a deterministic function of numpy's default_rng(seed) stream, so the draw order below
is the contract: nebula noise, (z, phi, accept) batches of star positions, masses, distances, the
`choice` of visible stars.  Output: (tex_h, tex_w, 3) float32 in [0, 1]."""
import os

import numpy as np
from PIL import Image

STAR_BRIGHTNESS = (0.03, 1.0)     # min, max
STAR_GAIN = 1.8
STAR_SATURATION = 0.3
STAR_SIGMA = (0.5, 1.7)           # Gaussian blob radius in texels
MILKY_WAY_GLOW = 0.10
GALACTIC_CENTER_GLOW = 0.08

GAL_INCL = np.radians(62.87)      # galactic plane vs. equator
GAL_RA = np.radians(266.4)        # galactic centre
GAL_DEC = np.radians(-28.9)


def blackbody_rgb(temp_k):
    """Tanner-Helland colour-temperature fit, vectorised; returns (..., 3) float32."""
    t = temp_k / 100.0
    hot = np.maximum(t - 60, 1e-6)
    red = np.where(t <= 66, 1.0, np.clip(1.292936 * np.power(hot, -0.1332047592), 0, 1))
    green = np.where(t <= 66, np.clip(0.390082 * np.log(np.maximum(t, 1e-6)) - 0.631841, 0, 1),
                     np.clip(1.129891 * np.power(hot, -0.0755148492), 0, 1))
    blue = np.where(t >= 66, 1.0, np.where(
        t <= 19, 0.0, np.clip(0.543207 * np.log(np.maximum(t - 10, 1e-6)) - 1.19625, 0, 1)))
    return np.stack([red, green, blue], axis=-1).astype(np.float32)


def _galactic_latitude(dec, ra):
    sin_b = np.sin(dec) * np.cos(GAL_INCL) - np.cos(dec) * np.sin(GAL_INCL) * np.sin(ra - GAL_RA)
    return np.arcsin(np.clip(sin_b, -1, 1))


def _star_positions(rng, n_stars):
    """Rejection-sample directions, denser towards the galactic plane and centre."""
    phis, thetas = [], []
    batch = n_stars * 3
    while len(phis) < n_stars:
        z = rng.uniform(-1, 1, batch)
        phi = rng.uniform(0, 2 * np.pi, batch)
        theta = np.arccos(np.clip(z, -1, 1))
        dec = np.pi / 2 - theta
        b = _galactic_latitude(dec, phi)
        prob = 0.15 + 0.85 * np.exp(-0.5 * (b / np.radians(8)) ** 2)
        cos_d = (np.sin(dec) * np.sin(GAL_DEC) + np.cos(dec) * np.cos(GAL_DEC) * np.cos(phi - GAL_RA))
        dist = np.arccos(np.clip(cos_d, -1, 1))
        prob += 0.3 * np.exp(-0.5 * (dist / np.radians(20)) ** 2)
        prob = prob / prob.max()
        keep = rng.random(batch) < prob
        need = n_stars - len(phis)
        phis.extend(phi[keep][:need])
        thetas.extend(theta[keep][:need])
    return np.array(phis[:n_stars]), np.array(thetas[:n_stars])


def _star_population(rng, n_stars):
    """Salpeter masses + exponential distances, magnitude-limited; returns (mass, apparent mag)."""
    alpha, m_lo, m_hi = 2.35, 0.08, 50.0
    n = n_stars * 30
    u = rng.random(n)
    mass = (m_lo ** (1 - alpha) + u * (m_hi ** (1 - alpha) - m_lo ** (1 - alpha))) ** (1 / (1 - alpha))
    expo = np.where(mass < 0.43, 2.3, np.where(mass < 2.0, 4.0, np.where(mass < 55.0, 3.5, 1.0)))
    abs_mag = -2.5 * np.log10(np.power(mass, expo) + 1e-30) + 4.83
    dist = np.clip(rng.exponential(scale=200.0, size=n), 1.0, 5000.0)
    app_mag = abs_mag + 5.0 * np.log10(dist / 10.0)
    visible = np.where(app_mag <= 8.0)[0]
    if len(visible) >= n_stars:
        pick = rng.choice(visible, size=n_stars, replace=False)
    else:
        pick = np.argsort(app_mag)[:n_stars]
    return mass[pick], app_mag[pick]


def _splat_stars(texture, cx, cy, brightness, sigma, colors):
    tex_h, tex_w = texture.shape[:2]
    offs = np.arange(-4, 5, dtype=np.float32)
    dy, dx = (g.ravel() for g in np.meshgrid(offs, offs, indexing="ij"))
    px = (cx[:, None] + dx[None, :]).astype(int) % tex_w
    py = (cy[:, None] + dy[None, :]).astype(int)
    d2 = dx[None, :] ** 2 + dy[None, :] ** 2
    vals = brightness[:, None] * np.exp(-d2 / (2 * sigma[:, None] ** 2))
    ok = (py >= 0) & (py < tex_h)
    cols = np.repeat(colors, len(dx), axis=0)[ok.ravel()]
    np.add.at(texture, (py[ok], px[ok]), cols * vals[ok][:, None])


def _milky_way(tex_h, tex_w):
    vv, uu = np.meshgrid(np.linspace(0, np.pi, tex_h), np.linspace(0, 2 * np.pi, tex_w), indexing="ij")
    dec = np.pi / 2 - vv
    b = _galactic_latitude(dec, uu)
    sin_l = np.cos(dec) * np.cos(GAL_INCL) * np.sin(uu - GAL_RA) + np.sin(dec) * np.sin(GAL_INCL)
    cos_l = np.cos(dec) * np.cos(uu - GAL_RA)
    lon = np.arctan2(sin_l, cos_l)
    glow = MILKY_WAY_GLOW * np.exp(-0.5 * (b / np.radians(6)) ** 2)
    glow += GALACTIC_CENTER_GLOW * np.exp(-0.5 * (lon ** 2 + b ** 2) / np.radians(15) ** 2)
    arms = 0.4 + 0.6 * (0.5 + 0.5 * np.cos(4 * lon + np.radians(30)))
    near_plane = np.exp(-0.5 * (b / np.radians(8)) ** 2)
    glow *= (1.0 - near_plane) + near_plane * arms
    return glow


def generate_skybox(tex_w=2048, tex_h=1024, seed=42, n_stars=6000):
    rng = np.random.default_rng(seed)
    texture = np.full((tex_h, tex_w, 3), 0.003, dtype=np.float32)
    # faint nebula: coarse noise, bilinearly upsampled through an 8-bit image
    coarse = rng.random((tex_h // 16, tex_w // 16, 3)).astype(np.float32) * 0.06
    up = Image.fromarray((coarse * 255).astype(np.uint8)).resize((tex_w, tex_h), Image.Resampling.BILINEAR)
    texture += np.array(up) / 255.0 * 0.04

    phi_s, theta_s = _star_positions(rng, n_stars)
    cx = (phi_s / (2 * np.pi) * tex_w).astype(np.float32)
    cy = (theta_s / np.pi * tex_h).astype(np.float32)
    mass, mag = _star_population(rng, n_stars)
    rel = (mag - mag.min()) / (mag.max() - mag.min() + 1e-30)
    b_min, b_max = STAR_BRIGHTNESS
    brightness = (b_max - (b_max - b_min) * rel).astype(np.float32)
    brightness = np.clip(brightness * STAR_GAIN, 0, 1)
    sigma = (STAR_SIGMA[0] + (STAR_SIGMA[1] - STAR_SIGMA[0]) * brightness).astype(np.float32)
    colors = blackbody_rgb(np.clip(5778.0 * np.power(mass, 0.57), 2000, 50000))
    colors = STAR_SATURATION * colors + (1 - STAR_SATURATION) * np.ones_like(colors)
    _splat_stars(texture, cx, cy, brightness, sigma, colors)

    texture += _milky_way(tex_h, tex_w)[:, :, None] * np.array([1.0, 0.95, 0.85])
    return np.clip(texture, 0, 1)


def load_or_generate_skybox(skybox_path, tex_w=2048, tex_h=1024, n_stars=6000):
    """Returns (texture, tex_h, tex_w); loads an image file if given, else generates."""
    if skybox_path and os.path.isfile(skybox_path):
        print(f"Loading skybox: {skybox_path}")
        texture = np.array(Image.open(skybox_path).convert("RGB"), dtype=np.float32) / 255.0
        tex_h, tex_w = texture.shape[:2]
    else:
        print("Generating procedural skybox..." if not skybox_path
              else f"Texture not found: {skybox_path}, generating procedural skybox...")
        texture = generate_skybox(tex_w=tex_w, tex_h=tex_h, n_stars=n_stars)
    return texture, tex_h, tex_w
