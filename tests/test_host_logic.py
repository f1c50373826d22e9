"""CPU tests of the host side: grid construction vs reference goldens, C-ABI exports."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import svox2_b200 as sv
from svox2_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "structure.npz")


@pytest.mark.parametrize("key,reso,zo,sph", [
    ("links_16_1_1", 16, True, True), ("links_16_0_1", 16, False, True),
    ("links_[8,12,10]_0_1", [8, 12, 10], False, True), ("links_8_1_0", 8, True, False)])
def test_links_construction_bit_exact(key, reso, zo, sph):
    z = np.load(GOLD)
    g = sv.SparseGrid(reso=reso, use_z_order=zo, use_sphere_bound=sph, device="cpu")
    assert np.array_equal(g.links.numpy(), z[key])
    assert int(g.capacity) == int(z[key + "_cap"])
    assert g.density_data.shape == (g.capacity, 1) and g.sh_data.shape == (g.capacity, 27)


def test_morton_sh_and_camera_rays():
    z = np.load(GOLD)
    assert np.array_equal(sv.utils.gen_morton(8).numpy(), z["morton8"])
    d = torch.from_numpy(z["sh_dirs"])
    for bd in (1, 4, 9):
        assert np.array_equal(sv.utils.eval_sh_bases(bd, d).numpy(), z[f"sh_bases_{bd}"])
    cam = sv.Camera(torch.from_numpy(z["cam_c2w"]), fx=40.0, fy=42.0, cx=15.5, cy=11.0, width=32, height=24)
    r = cam.gen_rays()
    assert np.array_equal(r.origins.numpy(), z["cam_origins"])
    assert np.abs(r.dirs.numpy() - z["cam_dirs"]).max() < 1e-7
    cam2 = sv.Camera(torch.from_numpy(z["ndc_c2w"]), fx=40.0, fy=42.0, cx=15.5, cy=11.0, width=32,
                     height=24, ndc_coeffs=(2 * 40.0 / 32, 2 * 42.0 / 24))
    r2 = cam2.gen_rays()
    assert np.abs(r2.origins.numpy() - z["ndc_origins"]).max() < 1e-6
    assert np.abs(r2.dirs.numpy() - z["ndc_dirs"]).max() < 1e-6


def test_world_grid_roundtrip_and_rand_cells():
    g = sv.SparseGrid(reso=[8, 12, 10], radius=[1.0, 0.8, 1.2], center=[0.1, 0.0, -0.2], device="cpu")
    p = torch.randn(20, 3)
    assert torch.allclose(g.grid2world(g.world2grid(p)), p, atol=1e-5)
    np.random.seed(3)
    cells = g._get_rand_cells(0.25)
    n = 8 * 12 * 10
    assert cells.dtype == torch.int32 and cells.numel() == int(0.25 * n)
    assert ((cells >= 0) & (cells < n)).all()
    d = (cells[1:] - cells[:-1]).numpy()
    assert ((d == 1) | (d == 1 - n)).all()          # contiguous run, wrapped once at most
    assert g._get_rand_cells(1.0) is None


def test_no_silent_cpu_fallback():
    g = sv.SparseGrid(reso=8, device="cpu")
    rays = sv.Rays(torch.zeros(4, 3), torch.ones(4, 3))
    with pytest.raises(NotImplementedError):
        g.volume_render(rays, use_kernel=False)
    with pytest.raises(NotImplementedError):
        g.volume_render(rays)
    with pytest.raises(NotImplementedError):
        g.sample(torch.zeros(3, 3))
    with pytest.raises(NotImplementedError):
        g.volume_render_fused(rays, torch.zeros(4, 3))
    with pytest.raises(NotImplementedError):
        sv.SparseGrid(reso=8, basis_type=sv.BASIS_TYPE_MLP)


def test_save_load_roundtrip(tmp_path):
    g = sv.SparseGrid(reso=8, basis_dim=4, use_sphere_bound=True, device="cpu")
    with torch.no_grad():
        g.density_data[:] = torch.rand_like(g.density_data)
        g.sh_data[:] = torch.randn_like(g.sh_data)
    path = str(tmp_path / "ckpt.npz")
    g.save(path)
    z = np.load(path)
    assert z["sh_data"].dtype == np.float16 and z["links"].dtype == np.int32
    g2 = sv.SparseGrid.load(path)
    assert torch.equal(g2.links, g.links) and g2.basis_dim == 4
    assert torch.equal(g2.density_data, g.density_data)
    assert torch.allclose(g2.sh_data, g.sh_data, atol=2e-3, rtol=1e-3)


def test_cabi_library_loads_and_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "svox2_b200.h")).read()
    declared = set(re.findall(r"SVOX2_API[^;(]*?\b(svox2_\w+)\s*\(", hdr))
    assert len(declared) >= 22
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in include/svox2_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), "ctypes table out of sync with the header"
    lib = _lib.lib()
    assert b"sm_100a" in lib.svox2_version()
    assert lib.svox2_render_workspace_bytes(5000) >= 5000 * 8
    # 128^3 pyramid: 64^3 + 32^3 + ... + 1
    assert lib.svox2_accel_workspace_bytes(128, 128, 128) == sum((128 >> k) ** 3 for k in range(1, 8))


def test_cabi_argument_errors_are_reported_not_printed():
    lib = _lib.lib()
    g = _lib.GridT()  # all-null grid
    o = _lib.RenderOptsT()
    st = lib.svox2_volume_render_cuvol(ctypes.byref(g), None, None, 0, ctypes.byref(o), None, None, 0,
                                       None, None)
    assert st == 1
    assert b"non-null" in lib.svox2_last_error()
    with pytest.raises(_lib.Svox2Error):
        _lib.check(st)


def test_cabi_extension_entry_points_validate_their_arguments():
    """The multi-GPU / slab extensions reject bad arguments before touching the GPU."""
    lib = _lib.lib()
    buf = (ctypes.c_float * 64)()
    rows = (ctypes.c_int64 * 2)(0, 1)
    p = ctypes.addressof
    # pack: stride not a multiple of 4, stride too small for the mask flags, null pointers
    assert lib.svox2_pack_grad_rows(p(buf), 28, 27, p(buf), 1, p(rows), 2, p(buf), 30, None, None, None) == 1
    assert b"pack_stride" in lib.svox2_last_error()
    assert lib.svox2_pack_grad_rows(p(buf), 28, 27, p(buf), 1, p(rows), 2, p(buf), 28, p(buf), None, None) == 1
    assert lib.svox2_pack_grad_rows(None, 28, 27, None, 1, p(rows), 2, p(buf), 28, None, None, None) == 1
    assert lib.svox2_pack_grad_rows(None, 28, 27, None, 1, None, 0, None, 28, None, None, None) == 0   # empty: ok
    assert lib.svox2_unpack_rows(p(buf), 28, p(rows), 2, p(buf), 27, 27, p(buf), 1, 1, None, None, None) == 1
    assert b"padded" in lib.svox2_last_error()
    assert lib.svox2_rmsprop_step_packed(None, None, 28, 27, None, None, 1, p(rows), 2, p(buf), 28, None,
                                         None, None, 0, None, 0, 0.9, 0.1, 0.1, 1e-8, -1e9, -1e9, None) == 1
    # slab: range outside the grid, rank outside the slab count
    g = _lib.GridT()
    o = _lib.RenderOptsT()
    assert lib.svox2_slab_trace(ctypes.byref(g), None, None, 0, ctypes.byref(o), 0, 4, None, None, 0, None) == 1
    assert lib.svox2_slab_prefix(p(buf), p(buf), 4, 2, 2, p(buf), None) == 1
    assert b"slab" in lib.svox2_last_error()
    assert lib.svox2_slab_combine(p(buf), p(buf), p(buf), p(buf), 4, 0, 0, 1.0, 0.0, p(buf), p(buf),
                                  p(buf), None) == 1
    assert lib.svox2_slab_prefix(None, None, 0, 1, 0, None, None) == 0                                 # empty: ok


@pytest.mark.parametrize("background", [False, True])
def test_checkpoint_interop_with_reference(ref_svox2, tmp_path, background):
    """ckpt.npz written by this package loads in the unmodified reference and vice versa
    (svox2.py:1526-1628): f32 density, f16 SH, int32 links, MSI arrays, radius/center."""
    kw = dict(reso=[10, 8, 12], basis_dim=4, use_sphere_bound=True, device="cpu", radius=[1.0, 0.8, 1.2],
              center=[0.1, 0.0, -0.2])
    if background:
        kw.update(background_nlayers=4, background_reso=8)
    ours = sv.SparseGrid(**kw)
    with torch.no_grad():
        ours.density_data[:] = torch.rand_like(ours.density_data)
        ours.sh_data[:] = torch.randn_like(ours.sh_data)
        if background:
            ours.background_data[:] = torch.rand_like(ours.background_data)
    p1 = str(tmp_path / "ours.npz")
    ours.save(p1)
    theirs = ref_svox2.SparseGrid.load(p1, device="cpu")
    assert torch.equal(theirs.links, ours.links)
    assert torch.equal(theirs.density_data.data, ours.density_data.data)
    assert torch.equal(theirs.sh_data.data, ours.sh_data.data.half().float())
    assert torch.allclose(theirs.radius, ours.radius) and torch.allclose(theirs.center, ours.center)
    if background:
        assert torch.equal(theirs.background_data.data, ours.background_data.data)
        assert torch.equal(theirs.background_links, ours.background_links)
    # and back: the reference's own file through our loader, bit-identical arrays
    p2 = str(tmp_path / "theirs.npz")
    theirs.save(p2)
    back = sv.SparseGrid.load(p2, device="cpu")
    z1, z2 = np.load(p1), np.load(p2)
    assert sorted(z1.files) == sorted(z2.files)
    for k in z1.files:
        assert z1[k].dtype == z2[k].dtype and np.array_equal(z1[k], z2[k]), k
    assert torch.equal(back.links, ours.links) and back.basis_dim == ours.basis_dim
    assert torch.equal(back.sh_data.data, theirs.sh_data.data)
    assert back.use_background == background


def test_slab_partition_host_logic():
    """SlabTrainer's partition of rows, halo planes and the TV cell run (no GPU work involved)."""
    from svox2_b200.slab import SlabTrainer, balanced_slabs
    g = sv.SparseGrid(reso=[24, 10, 12], basis_dim=1, use_sphere_bound=True, device="cpu")
    N = g.density_data.size(0)
    for G in (1, 2, 3, 5):
        bounds = balanced_slabs(g.links, G)
        assert bounds[0] == 0 and bounds[-1] == 24 and all(bounds[i] < bounds[i + 1] for i in range(G))
        trainers = [SlabTrainer(g, world_size=G, rank=r, bounds=bounds) for r in range(G)]
        excl = torch.stack([t_.excl_mask for t_ in trainers]).sum(0)
        assert torch.equal(excl, torch.ones(N, dtype=excl.dtype))           # every row has one owner
        for r in range(G - 1):                                               # neighbours agree on the plane
            assert torch.equal(trainers[r].halo_right, trainers[r + 1].halo_left)
            co_owned = trainers[r].own_mask & trainers[r + 1].own_mask
            assert int(co_owned.sum()) == trainers[r].halo_right.numel()
        assert trainers[0].halo_left is None and trainers[-1].halo_right is None
        # the wrapped TV run is split without gaps or overlaps
        total = 24 * 10 * 12
        for seed in range(4):
            cells = []
            for t_ in trainers:
                np.random.seed(seed)
                runs, n = t_._tv_runs()
                for a, c in runs:
                    cells.append(np.arange(a, a + c))
            np.random.seed(seed)
            start, n = g._draw_cell_run(trainers[0].tv_sparsity)
            want = np.sort((start + np.arange(n)) % total)
            assert np.array_equal(np.sort(np.concatenate(cells)), want)


def test_slab_compositing_algebra():
    """The identities svox2_slab_prefix / svox2_slab_combine rely on, in numpy:
    compositing a ray slab by slab with the entry log-transmittance added equals the sequential
    march, and the reference's backward `accum` (cuvol.cu:248-264) at a slab's first sample is
    rgb.gout minus gout.rgb_partial of the slabs crossed before it."""
    rng = np.random.default_rng(0)
    n, G = 40, 4
    p = rng.uniform(0.0, 0.3, n)                      # sigma * delta per sample
    col = rng.uniform(0.0, 1.0, (n, 3))               # clamped sample colours
    cuts = [0, 9, 17, 31, n]                          # samples of slab k: cuts[k]..cuts[k+1]
    bg, gt = 1.0, rng.uniform(0, 1, 3)
    # sequential (reference) forward
    lt, rgb, w = 0.0, np.zeros(3), np.zeros(n)
    for i in range(n):
        w[i] = np.exp(lt) * (1 - np.exp(-p[i]))
        rgb += w[i] * col[i]
        lt -= p[i]
    rgb_seq = rgb + np.exp(lt) * bg
    # slab-wise forward
    logT = np.array([-p[cuts[k]:cuts[k + 1]].sum() for k in range(G)])
    lt_in = np.concatenate([[0.0], np.cumsum(logT)[:-1]])
    part = np.zeros((G, 3))
    for k in range(G):
        local = 0.0
        for i in range(cuts[k], cuts[k + 1]):
            part[k] += np.exp(lt_in[k] + local) * (1 - np.exp(-p[i])) * col[i]
            local -= p[i]
    rgb_slab = part.sum(0) + np.exp(logT.sum()) * bg
    assert np.allclose(rgb_slab, rgb_seq, rtol=1e-12)
    # backward: accum before sample i = rgb.gout - sum_{j<i} w_j (col_j . gout)
    gout = (rgb_seq - gt) * (2.0 / 3.0)
    accum = rgb_seq @ gout - np.concatenate([[0.0], np.cumsum(w * (col @ gout))[:-1]])
    for k in range(G):
        start = rgb_seq @ gout - sum(part[j] @ gout for j in range(k))
        assert np.isclose(start, accum[cuts[k]], rtol=1e-10)


def test_compat_alias_exposes_importable_submodules():
    """`PYTHONPATH=compat:repo` must satisfy every `svox2` import of the reference's opt/ scripts:
    `import svox2.utils` (render_imgs.py), `from svox2.utils import convert_to_ndc`
    (util/llff_dataset.py), `svox2.csrc`.  Run in a fresh interpreter: the reference fixture of this
    session may already own the name `svox2`."""
    import subprocess
    import sys
    code = ("import svox2, svox2.utils, svox2.csrc\n"
            "from svox2.utils import convert_to_ndc, eval_sh_bases, Timing\n"
            "from svox2 import SparseGrid, Camera, Rays, RenderOptions\n"
            "assert svox2.utils.SH_C0 > 0.28 and callable(convert_to_ndc)\n"
            "assert hasattr(svox2.csrc, 'sample_grid') and svox2.BASIS_TYPE_MLP == 255\n"
            "assert svox2.SparseGrid.__module__.startswith('svox2_b200')\n")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "compat"), ROOT]))
    subprocess.run([sys.executable, "-c", code], check=True, env=env, cwd="/")
