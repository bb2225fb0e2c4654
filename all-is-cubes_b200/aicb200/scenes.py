"""Deterministic synthetic Spaces for tests and bench.

A stateless 64-bit hash makes the scenes toolchain independent:
    h(seed,x,y,z) = splitmix64(seed ^ x*0x9E3779B97F4A7C15 ^ y*0xC2B2AE3D27D4EB4F ^ z*0x165667B19E3779F9)
"""
from __future__ import annotations

import numpy as np

from . import (Block, Camera, GraphicsOptions, Space, Viewport, eye_for_look_at)

U64 = np.uint64
_K1, _K2, _K3 = U64(0x9E3779B97F4A7C15), U64(0xC2B2AE3D27D4EB4F), U64(0x165667B19E3779F9)


def splitmix64(x):
    with np.errstate(over="ignore"):
        x = (np.asarray(x, dtype=U64) + _K1).astype(U64)
        z = x
        z = ((z ^ (z >> U64(30))) * U64(0xBF58476D1CE4E5B9)).astype(U64)
        z = ((z ^ (z >> U64(27))) * U64(0x94D049BB133111EB)).astype(U64)
        return (z ^ (z >> U64(31))).astype(U64)


def hash3(seed, x, y, z):
    with np.errstate(over="ignore"):
        x = np.asarray(x).astype(np.int64).astype(U64)
        y = np.asarray(y).astype(np.int64).astype(U64)
        z = np.asarray(z).astype(np.int64).astype(U64)
        return splitmix64(U64(seed) ^ (x * _K1) ^ (y * _K2) ^ (z * _K3))


def grid_hash(seed, size):
    x, y, z = np.meshgrid(np.arange(size[0]), np.arange(size[1]), np.arange(size[2]), indexing="ij")
    return hash3(seed, x, y, z)


def _color_from_hash(h, alpha=1.0):
    h = int(h)
    return (((h >> 16) & 255) / 255.0, ((h >> 24) & 255) / 255.0, ((h >> 32) & 255) / 255.0, alpha)


def make_palette(seed, n, alpha=1.0, emissive_every=0):
    pal = np.zeros((n, 8), dtype=np.float32)
    hs = splitmix64(np.arange(n, dtype=U64) + U64(seed) * U64(1000003))
    for i in range(n):
        pal[i, :4] = _color_from_hash(hs[i], alpha)
        if emissive_every and i % emissive_every == emissive_every - 1:
            pal[i, 4:7] = (pal[i, 0] * 2.0, pal[i, 1] * 2.0, pal[i, 2] * 2.0)
    return pal


def make_voxel_block(seed, resolution=16, palette_size=16, alpha=1.0, fill_mask=3, partial_bounds=True,
                     transparent_palette_entry=False, emissive_every=0):
    """A recursive block: voxels solid iff h&fill_mask != 0 inside a hash-chosen sub-box
    (exercises voxel_bounds smaller than resolution^3, voxel_storage.rs:176-178)."""
    r = resolution
    hb = int(splitmix64(U64(seed)))
    if partial_bounds:
        lo = [(hb >> (4 * a)) % max(1, r // 4) for a in range(3)]
        hi = [r - ((hb >> (12 + 4 * a)) % max(1, r // 4)) for a in range(3)]
    else:
        lo, hi = [0, 0, 0], [r, r, r]
    size = [hi[a] - lo[a] for a in range(3)]
    x, y, z = np.meshgrid(np.arange(lo[0], hi[0]), np.arange(lo[1], hi[1]), np.arange(lo[2], hi[2]), indexing="ij")
    h = hash3(seed * 7919 + 13, x, y, z)
    pal = np.zeros((palette_size + 1, 8), dtype=np.float32)  # entry 0 = AIR voxel
    pal[1:] = make_palette(seed, palette_size, alpha, emissive_every)
    if transparent_palette_entry:
        pal[1, 3] = 0.25
        pal[2, 3] = 0.5
    solid = (h & U64(fill_mask)) != 0
    idx = np.where(solid, 1 + ((h >> U64(8)) % U64(palette_size)).astype(np.int64), 0).astype(np.uint16)
    assert idx.shape == tuple(size)
    return Block(resolution=r, voxel_lower=lo, indices=idx, palette=pal)


OCTANT_SKY = [  # Sky::Octants as in content/testing.rs:124-137 (a coloured test sky)
    (0.1, 0.1, 0.1), (0.1, 0.1, 0.4), (0.2, 0.2, 0.2), (0.2, 0.2, 0.8),
    (0.4, 0.1, 0.1), (0.4, 0.1, 0.4), (0.9, 0.9, 0.9), (0.8, 0.8, 1.0),
]


def noise_light(seed, block_ids, blocks, lo=96, hi=176):
    """A PackedLight volume: hash noise in [lo,hi] per channel; status Visible for cubes whose
    block is invisible/air, Opaque (value 0) otherwise (light/data.rs:31-46)."""
    size = block_ids.shape
    h = grid_hash(seed, size)
    span = U64(hi - lo + 1)
    light = np.zeros(size + (4,), dtype=np.uint8)
    for c in range(3):
        light[..., c] = (lo + ((h >> U64(8 * c + 3)) % span).astype(np.int64)).astype(np.uint8)
    opaque_block = np.array([(not b.is_air) and b.indices is None and b.palette[0, 3] == 1.0 for b in blocks])
    opq = opaque_block[block_ids]
    light[..., 3] = np.where(opq, 128, 255)
    light[opq, 0:3] = 0
    return light


# ---------------------------------------------------------------------------------------------
# BASELINE.json configs
# ---------------------------------------------------------------------------------------------
def config_c0(n=32, seed=1):
    """CPU config: n^3 Space of solid/empty resolution-1 blocks (12.5 % fill), 16 colours."""
    h = grid_hash(seed, (n, n, n))
    solid = (h & U64(7)) == 0
    ids = np.where(solid, 1 + ((h >> U64(8)) & U64(15)).astype(np.int64), 0).astype(np.uint16)
    pal = make_palette(seed, 16)
    blocks = [Block.air()] + [Block(color=tuple(pal[i, :4])) for i in range(16)]
    return Space((0, 0, 0), ids, blocks)


def config_c1(n=128, seed=2, n_voxel_blocks=256, with_light=False, resolution=16):
    """n^3, resolution-16 recursive blocks (opaque), ground slab of res-1 blocks for y <= n/4."""
    h = grid_hash(seed, (n, n, n))
    y = np.arange(n)[None, :, None]
    above = y > n // 4
    pal = make_palette(seed + 100, 16)
    ground_ids = 1 + ((h >> U64(8)) & U64(15)).astype(np.int64)
    voxel_ids = 17 + ((h >> U64(8)) % U64(n_voxel_blocks)).astype(np.int64)
    ids = np.where(above, np.where((h & U64(15)) == 0, voxel_ids, 0), ground_ids).astype(np.uint16)
    blocks = [Block.air()] + [Block(color=tuple(pal[i, :4])) for i in range(16)]
    blocks += [make_voxel_block(seed * 1000 + i, resolution=resolution) for i in range(n_voxel_blocks)]
    light = noise_light(seed + 5, ids, blocks) if with_light else None
    return Space((0, 0, 0), ids, blocks, light=light)


def config_c2(n=256, seed=3, n_voxel_blocks=64, with_light=False):
    """n^3 mixed transparent blocks: fill 1/8; 40 % of non-AIR cubes are alpha in
    {0.125,0.25,0.5} (res-1 and res-16), the rest opaque (deep accumulation)."""
    h = grid_hash(seed, (n, n, n))
    filled = (h & U64(7)) == 0
    sel = ((h >> U64(8)) % U64(100)).astype(np.int64)
    which = ((h >> U64(20)) & U64(0xffff)).astype(np.int64)
    blocks = [Block.air()]
    pal = make_palette(seed + 100, 16)
    opaque_single = list(range(len(blocks), len(blocks) + 16))
    blocks += [Block(color=tuple(pal[i, :4])) for i in range(16)]
    trans_single = []
    for a in (0.125, 0.25, 0.5):
        palt = make_palette(seed + 200 + int(a * 1000), 8, alpha=a)
        trans_single += list(range(len(blocks), len(blocks) + 8))
        blocks += [Block(color=tuple(palt[i, :4])) for i in range(8)]
    trans_voxel = list(range(len(blocks), len(blocks) + n_voxel_blocks))
    blocks += [make_voxel_block(seed * 1000 + i, alpha=(0.125, 0.25, 0.5)[i % 3]) for i in range(n_voxel_blocks)]
    opaque_voxel = list(range(len(blocks), len(blocks) + n_voxel_blocks))
    blocks += [make_voxel_block(seed * 2000 + i) for i in range(n_voxel_blocks)]
    opaque_single, trans_single = np.array(opaque_single), np.array(trans_single)
    trans_voxel, opaque_voxel = np.array(trans_voxel), np.array(opaque_voxel)
    ids = np.zeros((n, n, n), dtype=np.int64)
    ids = np.where(filled & (sel < 25), trans_single[which % len(trans_single)], ids)
    ids = np.where(filled & (sel >= 25) & (sel < 40), trans_voxel[which % len(trans_voxel)], ids)
    ids = np.where(filled & (sel >= 40) & (sel < 85), opaque_single[which % len(opaque_single)], ids)
    ids = np.where(filled & (sel >= 85), opaque_voxel[which % len(opaque_voxel)], ids)
    ids = ids.astype(np.uint16)
    light = noise_light(seed + 5, ids, blocks) if with_light else None
    return Space((0, 0, 0), ids, blocks, light=light)


def config_c4(n=256, seed=4):
    """BASELINE configs[4]: the C1 layout with res-1 blocks — ground slab of opaque blocks for
    y < n/4, 1/16 of the cubes above it opaque (one in fifteen of them an emitter) — LightPhysics::Rays{30},
    Sky::Octants as content/testing.rs:124-137, light all NO_RAYS (to be converged by the light kernels)."""
    h = grid_hash(seed, (n, n, n))
    pal = make_palette(seed, 14)
    blocks = [Block.air()] + [Block(color=tuple(pal[i, :4])) for i in range(14)] + \
             [Block(color=(0.1, 0.1, 0.1, 1.0), emission=(4.0, 3.5, 2.0))]
    ids = np.where((h & U64(15)) == 0, 1 + ((h >> U64(8)) % U64(15)).astype(np.int64), 0).astype(np.uint16)
    ids[:, : n // 4, :] = 1 + ((h[:, : n // 4, :] >> U64(8)) % U64(14)).astype(np.uint16)
    light = np.zeros((n, n, n, 4), dtype=np.uint8)
    light[..., 3] = 1
    return Space((0, 0, 0), ids, blocks, light=light, sky_colors=OCTANT_SKY, light_max_distance=30)


def c4_edits(space, n_edits, batch):
    """The `batch`-th set of random block edits of C4: cubes from the ground surface upwards, any block of the table."""
    n = space.size[0]
    rng = np.random.default_rng(7 + batch)
    cubes = np.stack([rng.integers(0, n, n_edits), rng.integers(n // 4 - 2, n, n_edits), rng.integers(0, n, n_edits)],
                     axis=1).astype(np.int32)
    return cubes, rng.integers(0, len(space.blocks), n_edits).astype(np.uint16)


def small_mixed_scene(n=12, seed=7, lower=(-3, 2, -5), with_light=True, octant_sky=True):
    """A small Space exercising every code path: AIR, opaque / transparent / emissive /
    invisible-but-not-AIR single blocks, recursive blocks (res 2..16, partial bounds,
    transparent + emissive voxels), a light volume, an octant sky, a non-zero lower bound."""
    h = grid_hash(seed, (n, n, n))
    blocks = [Block.air()]
    pal = make_palette(seed, 6)
    blocks += [Block(color=tuple(pal[i, :4])) for i in range(4)]                       # 1-4 opaque
    blocks += [Block(color=(0.9, 0.2, 0.1, 0.5)), Block(color=(0.1, 0.8, 0.3, 0.125))]  # 5,6 transparent
    blocks += [Block(color=(0.2, 0.2, 0.2, 1.0), emission=(1.5, 0.5, 0.1))]            # 7 emissive opaque
    blocks += [Block(color=(0.0, 0.0, 0.0, 0.0), emission=(0.1, 0.3, 0.9))]            # 8 emissive, alpha 0
    blocks += [Block(color=(0.5, 0.5, 0.5, 0.0))]                                      # 9 invisible, not AIR
    blocks += [make_voxel_block(seed * 10 + 1, resolution=16, palette_size=6)]         # 10
    blocks += [make_voxel_block(seed * 10 + 2, resolution=8, palette_size=4, transparent_palette_entry=True)]  # 11
    blocks += [make_voxel_block(seed * 10 + 3, resolution=4, palette_size=3, alpha=0.5, emissive_every=2)]     # 12
    blocks += [make_voxel_block(seed * 10 + 4, resolution=2, palette_size=2, partial_bounds=False)]            # 13
    blocks += [make_voxel_block(seed * 10 + 5, resolution=16, palette_size=5, fill_mask=7)]                    # 14
    sel = (h % U64(64)).astype(np.int64)
    ids = np.zeros((n, n, n), dtype=np.int64)
    ids = np.where(sel < 14, sel + 1, ids)  # ~22 % filled, every block type present
    ids = np.where(ids > 14, 0, ids).astype(np.uint16)
    light = noise_light(seed + 5, ids, blocks) if with_light else None
    return Space(lower, ids, blocks, light=light, sky_colors=OCTANT_SKY if octant_sky else None)


def standard_camera(space: Space, options: GraphicsOptions, width: int, height: int, direction=(1.0, 0.6, 1.0),
                    distance_scale=1.0) -> Camera:
    """Camera::new(options, Viewport::with_scale(1, [w,h])) + look_at_y_up(eye_for_look_at(bounds, dir),
    bounds.center())."""
    cam = Camera(options, Viewport.with_scale(1.0, (width, height)))
    center = [space.lower[a] + space.size[a] / 2.0 for a in range(3)]
    eye = eye_for_look_at(space.lower, space.size, direction)
    if distance_scale != 1.0:
        eye = np.array(center) + (eye - np.array(center)) * distance_scale
    cam.look_at_y_up(eye, center)
    return cam
