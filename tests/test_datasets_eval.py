"""dataset loaders + ray generation (NSVF / T&T, LLFF-NDC) and the SSIM of
the evaluation loop, compared with the reference's OWN loaders (oracle/_ref/opt/util/*, unmodified)
on generated scenes in both on-disk formats.  CPU; the device-side versions run the same torch code.
`imageio` is absent from this image: tests/stubs/imageio (PIL-backed) stands in for it."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
STUBS = os.path.join(ROOT, "tests", "stubs")


@pytest.fixture(scope="module")
def ref_util(ref_svox2):
    if not os.path.isfile(os.path.join(REF, "opt", "util", "util.py")):
        pytest.skip("oracle/_ref/opt not installed (oracle/build_ref.sh)")
    for p in (STUBS, os.path.join(REF, "opt")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import util.util as U           # noqa: E402  (the reference's opt/util package)
    import util.nsvf_dataset as N   # noqa: E402
    import util.llff_dataset as L   # noqa: E402
    return dict(util=U, nsvf=N, llff=L)


def _look_at(pos, target, rng):
    fwd = target - pos
    fwd /= np.linalg.norm(fwd)
    up = np.array([0.0, 0.0, 1.0]) + rng.normal(0, 0.05, 3)
    right = np.cross(fwd, up)
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    m = np.eye(4)
    m[:3, :3] = np.stack([right, down, fwd], 1)   # OpenCV: x right, y down, z forward
    m[:3, 3] = pos
    return m


def _png(path, rng, h, w, alpha=False):
    from PIL import Image
    arr = rng.integers(0, 256, (h, w, 4 if alpha else 3), dtype=np.uint8)
    Image.fromarray(arr).save(path)


def _write_nsvf(root, rng, n=20, h=18, w=26, alpha=True, plain_intrinsics=False):
    os.makedirs(os.path.join(root, "rgb"))
    os.makedirs(os.path.join(root, "pose"))
    for i in range(n):
        prefix = "0" if i % 5 else ("1" if i % 10 else "2")
        name = f"{prefix}_{i:04d}"
        pos = rng.normal(0, 1, 3)
        pos = pos / np.linalg.norm(pos) * rng.uniform(3.0, 4.5) + np.array([0.3, -0.2, 0.5])
        np.savetxt(os.path.join(root, "pose", name + ".txt"), _look_at(pos, rng.normal(0, 0.2, 3), rng))
        _png(os.path.join(root, "rgb", name + ".png"), rng, h, w, alpha)
    with open(os.path.join(root, "intrinsics.txt"), "w") as f:
        if plain_intrinsics:
            f.write(f"31.5 {w / 2} {h / 2} 0.\n0. 0. 0.\n0.\n1.\n{h} {w}\n")
        else:
            f.write(f"30.0 0.0 {w / 2 - 0.5} 0.0\n0.0 32.0 {h / 2 + 0.25} 0.0\n0.0 0.0 1.0 0.0\n0.0 0.0 0.0 1.0\n")


def _write_llff(root, rng, n=17, h=24, w=32, prescaled=False):
    os.makedirs(os.path.join(root, "images"))
    rows = []
    for i in range(n):
        pos = np.array([rng.uniform(-1, 1), rng.uniform(-0.6, 0.6), rng.uniform(-0.1, 0.1)])
        c2w = _look_at(pos, np.array([0.0, 0.0, 6.0]) + rng.normal(0, 0.3, 3), rng)  # OpenCV
        # LLFF stores [down, right, backwards] columns + position, then a [H, W, f] column
        llff = np.stack([c2w[:3, 1], c2w[:3, 0], -c2w[:3, 2], c2w[:3, 3], np.array([h, w, 27.5])], 1)
        rows.append(np.concatenate([llff.reshape(-1), [rng.uniform(2.0, 3.0), rng.uniform(20, 30)]]))
        _png(os.path.join(root, "images", f"img_{i:03d}.png"), rng, h, w)
    np.save(os.path.join(root, "poses_bounds.npy"), np.stack(rows))
    if prescaled:
        os.makedirs(os.path.join(root, "images_2"))
        for i in range(n):
            _png(os.path.join(root, "images_2", f"img_{i:03d}.png"), rng, h // 2, w // 2)


def _same_rays(ours, ref_rays, tol=2e-6):
    rays, gt = ours
    assert rays.origins.shape == ref_rays.origins.shape
    assert (rays.origins - ref_rays.origins).abs().max().item() <= tol
    assert (rays.dirs - ref_rays.dirs).abs().max().item() <= tol
    assert torch.equal(gt, ref_rays.gt)


@pytest.mark.parametrize("alpha,plain,scale", [(True, False, 1.0), (False, True, 1.0), (True, False, 0.5)])
def test_nsvf_loader_and_rays_match_reference_loader(ref_util, tmp_path, alpha, plain, scale):
    from svox2_b200 import datasets as D
    rng = np.random.default_rng(3)
    root = str(tmp_path / "nsvf")
    _write_nsvf(root, rng, alpha=alpha, plain_intrinsics=plain)
    for split in ("train", "test"):
        ref = ref_util["nsvf"].NSVFDataset(root, split, scale=scale)
        ours = D.load_nsvf(root, split, device="cpu", scale=scale)
        assert ours.n_images == ref.n_images and ours.names == ref.img_files
        assert torch.equal(ours.gt, ref.gt)
        assert (ours.c2w - ref.c2w).abs().max().item() < 1e-6
        assert abs(ours.scene_scale - ref.scene_scale) < 1e-9 * abs(ref.scene_scale) + 1e-12
        for f in ("fx", "fy", "cx", "cy"):
            assert abs(getattr(ours.intrins, f) - getattr(ref.intrins_full, f)) < 1e-9
        if split == "train":
            _same_rays(ours.gen_rays(), ref.rays_init)
            ref.gen_rays(factor=2)
            _same_rays(ours.gen_rays(factor=2), ref.rays_init)
    assert isinstance(D.auto_dataset(root, "val", device="cpu", scale=scale), D.PosedImageSet)


@pytest.mark.parametrize("scale,prescaled", [(0.5, False), (0.25, False), (0.5, True)])  # (the reference loader finds no images at scale 1)
def test_llff_loader_and_ndc_rays_match_reference_loader(ref_util, tmp_path, scale, prescaled):
    from svox2_b200 import datasets as D
    rng = np.random.default_rng(4)
    root = str(tmp_path / "llff")
    _write_llff(root, rng, prescaled=prescaled)
    for split in ("train", "test"):
        ref = ref_util["llff"].LLFFDataset(root, split, scale=scale)
        ours = D.load_llff(root, split, device="cpu", scale=scale)
        assert ours.n_images == ref.n_images
        assert torch.equal(ours.gt, ref.gt)
        assert (ours.c2w - ref.c2w).abs().max().item() < 1e-5
        for f in ("fx", "fy", "cx", "cy"):
            assert abs(getattr(ours.intrins, f) - getattr(ref.intrins_full, f)) < 1e-9
        assert np.allclose(ours.ndc_coeffs, ref.ndc_coeffs, rtol=0, atol=1e-12)
        assert np.allclose(ours.scene_radius, ref.scene_radius) and not ours.use_sphere_bound
        if split == "train":
            # same poses into both ray generators: isolates the ray arithmetic from the pose fit
            ours.c2w = ref.c2w.clone()
            _same_rays(ours.gen_rays(), ref.rays_init, tol=5e-6)
            assert (ours.gen_rays()[0].dirs.norm(dim=-1) - 1).abs().max().item() < 1e-5
    cam = ours.camera(0)
    assert cam.ndc_coeffs[0] == pytest.approx(ours.ndc_coeffs[0])
    assert isinstance(D.auto_dataset(root, "test", device="cpu", scale=scale), D.PosedImageSet)


def test_ssim_matches_reference_function(ref_util):
    from svox2_b200.evaluate import compute_ssim
    g = torch.Generator().manual_seed(0)
    a = torch.rand(37, 52, 3, generator=g)
    b = (a + 0.1 * torch.randn(37, 52, 3, generator=g)).clamp(0, 1)
    want = ref_util["util"].compute_ssim(a, b)
    got = compute_ssim(a, b)
    assert abs(got.item() - want.item()) < 1e-6
    assert abs(compute_ssim(a, a).item() - 1.0) < 1e-6
    wm = ref_util["util"].compute_ssim(a, b, return_map=True)
    gm = compute_ssim(a, b, return_map=True)
    assert torch.allclose(gm, wm, atol=1e-6)
    batch = torch.stack([a, b])
    assert torch.allclose(compute_ssim(batch, torch.stack([b, b])),
                          ref_util["util"].compute_ssim(batch, torch.stack([b, b])), atol=1e-6)


def test_similarity_transform_matches_reference(ref_util):
    from svox2_b200.datasets import similarity_from_cameras
    rng = np.random.default_rng(9)
    poses = np.stack([_look_at(rng.normal(0, 2, 3) + np.array([0, 0, 3.0]), rng.normal(0, 0.3, 3), rng)
                      for _ in range(25)])
    T0, s0 = ref_util["util"].similarity_from_cameras(poses)
    T1, s1 = similarity_from_cameras(poses)
    assert np.allclose(T0, T1, atol=1e-12) and abs(s0 - s1) < 1e-12
