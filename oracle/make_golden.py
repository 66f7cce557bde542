"""Generate tests/golden/*.npz by running the UNMODIFIED reference (oracle/ref_loader.py).

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python -m oracle.make_golden

Every output array below is produced by the reference's own ``ArrayPSF``,
``ArrayPSFTransform.construct`` and ``ArrayPSFTransform.apply``; inputs come from seeded
generators in ``oracle/cpu_oracle.py``.  The fixtures are small so they can be committed.
"""
from __future__ import annotations

import os

import numpy as np

from oracle import cpu_oracle as o
from oracle import ref_loader

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _run(ref, name, shape, size, alpha, epsilon, *, seed, coords=None, src_kind="coma", psf_dtype=np.float64,
         image_dtype=np.float32, apply_kwargs=None, keep_kernel=False):
    apply_kwargs = apply_kwargs or {}
    if coords is None:
        coords = [tuple(int(v) for v in c) for c in o.covering(shape, size)]
    n = len(coords)
    if src_kind == "coma":
        src = o.coma_psf_cube(coords, size, shape, dtype=psf_dtype)
        tgt = o.gaussian_psf_cube(n, size, 3.0, dtype=psf_dtype)
    elif src_kind == "gauss43":
        src = o.gaussian_psf_cube(n, size, 4.0, dtype=psf_dtype)
        tgt = o.gaussian_psf_cube(n, size, 3.0, dtype=psf_dtype)
    elif src_kind == "identity":
        src = o.gaussian_psf_cube(n, size, 3.0, dtype=psf_dtype)
        tgt = src
    else:
        raise ValueError(src_kind)
    source = ref.psf.ArrayPSF(ref.util.IndexedCube(coords, src))
    target = source if tgt is src else ref.psf.ArrayPSF(ref.util.IndexedCube(coords, tgt))
    transform = ref.transform.ArrayPSFTransform.construct(source, target, alpha, epsilon)
    image = o.starfield(shape, seed=seed)
    if np.issubdtype(image_dtype, np.integer):
        image = np.clip(image, 0, np.iinfo(image_dtype).max)
    image = image.astype(image_dtype)
    out = np.ascontiguousarray(transform.apply(image, **apply_kwargs))
    payload = dict(
        image=image, coords=np.array(coords, dtype=np.int64), size=np.int64(size),
        source=src, target=tgt, alpha=np.float64(alpha), epsilon=np.float64(epsilon), out=out,
        apply_kwargs=np.array(repr(apply_kwargs)),
    )
    if keep_kernel:
        payload["kernel"] = transform._transfer_kernel.values
        payload["source_fft"] = source.fft_evaluations
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **payload)
    print(f"{name}: N={n} P={size} image={shape} out max={np.nanmax(np.abs(out)):.4g} "
          f"nan={int(np.isnan(out).sum())}")


def _builder_case(name="c5_builder_p32", size=32):
    """Config 5: ArrayPSFBuilder star-cutout PSFs (reference builder, detections from oracle/fake_sep.py)
    -> construct over an alpha/epsilon sweep -> apply."""
    import warnings

    ref = ref_loader.load_builder()
    frames, mask = o.builder_frames()
    payload = dict(n_frames=np.int64(len(frames)), shape=np.array(frames.shape[1:]), size=np.int64(size),
                   frames_checksum=np.float64(frames.sum()), mask_checksum=np.int64(mask.sum()))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for method, pct in (("median", 50), ("mean", 50), ("percentile", 30)):
            model, counts, patches = ref.builder.ArrayPSFBuilder(size).build(
                frames, num_workers=1, average_method=method, percentile=pct, image_mask=mask, return_patches=True)
            payload[f"values_{method}"] = model.values
            payload["coords"] = np.array(model.coordinates, dtype=np.int64)
            payload["counts"] = np.array([counts[tuple(c)] for c in ref.util.calculate_covering(frames.shape[1:], size)])
            payload["n_cutouts"] = np.int64(len(patches))
            payload["cutout_nans"] = np.int64(sum(int(np.isnan(p).sum()) for p in patches.values()))
        coords = [tuple(int(v) for v in c) for c in payload["coords"]]
        source = ref.psf.ArrayPSF(ref.util.IndexedCube(coords, payload["values_median"]))
        target = ref.psf.ArrayPSF(ref.util.IndexedCube(coords, o.gaussian_psf_cube(len(coords), size, 3.0)))
        image = frames[0]
        crop = (slice(64, 192), slice(64, 192))
        sweep = [(0.5, 0.3), (1.0, 0.1), (2.0, 0.05), (3.0, 0.01)]
        for k, (alpha, epsilon) in enumerate(sweep):
            out = ref.transform.ArrayPSFTransform.construct(source, target, alpha, epsilon).apply(image)
            payload[f"out_crop_{k}"] = np.ascontiguousarray(out[crop])
        payload["sweep"] = np.array(sweep)
        payload["crop"] = np.array([64, 192, 64, 192])
    np.savez_compressed(os.path.join(OUT, "builder", f"{name}.npz"), **payload)
    print(f"{name}: {int(payload['n_cutouts'])} cutouts ({int(payload['cutout_nans'])} NaN pixels), counts "
          f"{payload['counts'].min()}..{payload['counts'].max()}, core pixels "
          f"{(payload['values_median'] != 0).sum(axis=(1, 2)).min()}..{(payload['values_median'] != 0).sum(axis=(1, 2)).max()}")


def main():
    os.makedirs(os.path.join(OUT, "builder"), exist_ok=True)
    _builder_case()
    ref = ref_loader.load()
    os.makedirs(OUT, exist_ok=True)
    _run(ref, "p16_coma_a1", (48, 40), 16, 1.0, 0.1, seed=11, keep_kernel=True)
    _run(ref, "p32_coma_a3", (96, 80), 32, 3.0, 0.1, seed=12)
    _run(ref, "p32_gauss43_f32psf", (64, 64), 32, 3.0, 0.1, seed=13, src_kind="gauss43", psf_dtype=np.float32,
         keep_kernel=True)
    _run(ref, "p32_identity_f32psf", (64, 96), 32, 3.0, 0.1, seed=14, src_kind="identity", psf_dtype=np.float32)
    for mode in ("reflect", "constant", "edge", "wrap"):
        _run(ref, f"p16_pad_{mode}", (40, 56), 16, 1.0, 0.1, seed=15, apply_kwargs={"pad_mode": mode})
    _run(ref, "p16_tiny_image", (20, 20), 16, 1.0, 0.1, seed=16)          # pad reach > image: reflection iterates
    _run(ref, "p32_uint16", (64, 64), 32, 1.0, 0.05, seed=17, image_dtype=np.uint16)
    _run(ref, "p32_float64_image", (64, 64), 32, 2.0, 0.3, seed=18, image_dtype=np.float64)
    _run(ref, "p32_saturation", (96, 96), 32, 1.0, 0.1, seed=19, apply_kwargs={"saturation_threshold": 3000.0})
    _run(ref, "p32_saturation_dil2", (96, 96), 32, 1.0, 0.1, seed=20,
         apply_kwargs={"saturation_threshold": 2000.0, "saturation_dilation": 2, "neighborhood_width": 9})
    # arbitrary user coordinate list: irregular overlaps, holes, partly off-image corners
    rng = np.random.default_rng(21)
    odd = sorted({(int(r), int(c)) for r, c in zip(rng.integers(-20, 70, 30), rng.integers(-20, 70, 30))})
    _run(ref, "p16_irregular_coords", (64, 64), 16, 1.0, 0.1, seed=21, coords=odd)
    _run(ref, "p64_coma_a05", (128, 192), 64, 0.5, 0.01, seed=22)


if __name__ == "__main__":
    main()
