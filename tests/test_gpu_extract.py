"""Stack-less loading: SPA box cutting on the device against a line-by-line restatement of the
reference's loop (src/pyp/extract/core.py:452-500, its edge rule included), and micrograph / tilt series -> boxes ->
scorer without a stack on the host."""
import math

import numpy as np
import pytest

from common import refine_cfg, small_case

pytestmark = pytest.mark.gpu


def reference_boxes(image, boxes, boxsize, coordinate_binning=1):
    """extract_particles_non_mpi's cutting loop (extract/core.py:452-500), frames == 1, before fix_empty / normalisation."""
    nx, ny = image.shape[-2], image.shape[-1]
    out = []
    for box in boxes:
        minx = miny = 0
        maxx = maxy = boxsize
        minX = math.floor(box[1] / float(coordinate_binning) - math.floor(boxsize / 2.0))
        maxX = minX + boxsize
        minY = math.floor(box[0] / float(coordinate_binning) - math.floor(boxsize / 2.0))
        maxY = minY + boxsize
        if minX < 0:
            minx = -minX
            minX = 0
        elif maxX >= nx:
            maxx = -(maxX - nx + 1)
            maxX = nx - 1
        if minY < 0:
            miny = -minY
            minY = 0
        elif maxY >= ny:
            maxy = -(maxY - ny + 1)
            maxY = ny - 1
        inside = np.squeeze(image[int(minX):int(maxX), int(minY):int(maxY)]) if maxX > minX and maxY > minY else np.zeros((0, 0))
        if inside.ndim == 2 and min(inside.shape) > 0:
            raw = inside.mean() * np.ones([boxsize, boxsize])
            raw[int(minx):int(maxx), int(miny):int(maxy)] = inside
        else:
            raw = np.zeros([boxsize, boxsize])
        out.append(raw.astype(np.float32))
    return np.stack(out)


def test_spa_extraction_equals_the_reference_loop(engine):
    rng = np.random.default_rng(0)
    ny, nx, box = 300, 420, 64
    mic = rng.normal(10.0, 2.0, (ny, nx)).astype(np.float32)
    xy = np.array([[200.3, 150.7], [31.0, 31.9], [10.2, 100.0], [415.0, 40.0], [388.0, 268.0], [100.0, 290.5], [5.0, 4.0], [800.0, 100.0],
                   [388.0, 100.0], [64.0, 268.0]], dtype=np.float32)   # inside, touching, clipped on every side, outside, exactly reaching the edge
    for cbin in (1, 2):
        got = engine.spa_extract(mic, xy, box, cbin)
        want = reference_boxes(mic, xy.tolist(), box, cbin)
        copied = np.isclose(got, want, rtol=0, atol=0)
        # copied pixels are bit-identical; the padding value is the inside mean (float32 block sum here, numpy pairwise there)
        assert np.abs(got - want).max() <= 2e-5 * np.abs(want).max(), cbin
        assert copied.mean() > 0.8
    # a box that ends exactly at the micrograph's edge loses its last line, as in the reference (maxX >= nx)
    edge = engine.spa_extract(mic, np.array([[388.0, 100.0]], np.float32), box, 1)[0]
    assert np.all(edge[:, -1] == edge[0, -1]) and not np.all(edge[:, -2] == edge[0, -2])
    # an empty (constant) region becomes reproducible unit white noise (image.py:461-471)
    flat = np.full((ny, nx), 3.0, np.float32)
    a = engine.spa_extract(flat, xy[:1], box)[0]
    b = engine.spa_extract(flat, xy[:1], box)[0]
    assert np.array_equal(a, b) and abs(a.mean()) < 0.1 and 0.9 < a.std() < 1.1


def test_micrograph_to_scorer_without_a_host_stack(engine, oracle):
    """Boxes cut on the device and handed to load_images as a device tensor score exactly like the same boxes loaded from a
    host stack; same for the tilt-series extraction of csp mode -2."""
    import torch

    n, px = 64, 1.35
    ph, vol, rows, stack = small_case(n=n, n_part=12, snr=0.5)
    # paste the particles into a micrograph on a grid, keep their centres
    ny, nx = 3 * n + 40, 4 * n + 40
    mic = np.random.default_rng(1).normal(0, 1, (ny, nx)).astype(np.float32)
    xy = []
    for k in range(rows.size):
        cx, cy = 20 + n // 2 + (k % 4) * n, 20 + n // 2 + (k // 4) * n
        mic[cy - n // 2:cy + n // 2, cx - n // 2:cx + n // 2] = stack[k]
        xy.append((cx, cy))
    cfg = refine_cfg(n, px)
    engine.refine_configure(cfg)
    engine.set_reference(vol)
    engine.load_images(stack)
    want = engine.score(rows)
    dev = engine.spa_extract(mic, np.array(xy, np.float32), n, to_device=True)
    assert isinstance(dev, torch.Tensor) and dev.is_cuda and np.array_equal(dev.cpu().numpy(), stack)
    engine.refine_configure(cfg)
    engine.set_reference(vol)
    engine.load_images(dev)
    assert np.array_equal(engine.score(rows), want)
    # tilt series on the device -> boxes on the device (csp mode -2 without the stack file)
    series = torch.from_numpy(mic[None]).cuda()
    r2 = rows.copy()
    r2["imind"] = 0
    r2["original_x"], r2["original_y"] = np.array(xy, np.float32).T
    dev2 = engine.csp_extract(series, r2, n, 1)
    assert dev2.is_cuda and np.array_equal(dev2.cpu().numpy(), stack)
