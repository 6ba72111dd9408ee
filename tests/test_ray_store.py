"""Preprocessed-ray format: the streaming writer against files written by the live reference
(tests/golden/make_ray_store_golden.py), the loader against the reference's ``load_preprocessed_data``."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR

from inf_b200 import ray_store

STEMS = ("face_idxs", "vids_of_hit_faces", "barycentric_coords", "expected_rgbs", "unit_ray_dirs")


@pytest.fixture(scope="module")
def golden():
    z = np.load(os.path.join(GOLDEN_DIR, "ray_store.npz"))
    return {k: z[k] for k in z.files}


def _replay_intersector(golden):
    """The fake intersector's recorded outputs, view by view (ray_mesh_intersect's return convention)."""
    state = {"v": 0}

    def intersect(ray_origins, unit_ray_dirs, return_depth=False, camCv2world=None):
        v = state["v"]
        state["v"] += 1
        out = [torch.from_numpy(golden[f"hit{v}_{n}"]) for n in ("vids", "bary", "hit_ray_idxs", "face_idxs")]
        assert int(out[2].max()) < ray_origins.shape[0] == unit_ray_dirs.shape[0]
        if return_depth:
            assert camCv2world is not None
            out.append(torch.from_numpy(golden[f"hit{v}_hit_depth"]))
        return tuple(out)
    return intersect


def _write_views(golden, out_dir):
    H, W = int(golden["H"]), int(golden["W"])
    with ray_store.RayCacheWriter(out_dir, intersect=_replay_intersector(golden)) as w:
        kept = []
        for v in range(int(golden["n_views"])):
            depth = golden.get(f"depth{v}")
            kept.append(w.cache_single_view(torch.from_numpy(golden[f"cam{v}"]), torch.from_numpy(golden[f"K{v}"]),
                                            torch.from_numpy(golden[f"mask{v}"]).reshape(H, W),
                                            torch.from_numpy(golden[f"img{v}"]), depth_check=depth))
    return w, kept


def test_streaming_writer_files_are_byte_identical_to_the_reference(golden, tmp_path):
    w, kept = _write_views(golden, str(tmp_path))
    assert w.num_hits == sum(kept) == golden["loaded_expected_rgbs"].shape[0]
    assert kept[1] < golden["hit1_vids"].shape[0]                   # the depth check removed hits in view 1
    for stem in STEMS:
        with open(tmp_path / (stem + ".npy"), "rb") as fh:
            mine = np.frombuffer(fh.read(), dtype=np.uint8)
        assert np.array_equal(mine, golden["file_" + stem]), stem


def test_pixel_rays_match_reference_directions(golden):
    # the unit_ray_dirs the reference stored are its create_ray_origins_and_directions output, selected by hit index
    H, W = int(golden["H"]), int(golden["W"])
    cam, K = torch.from_numpy(golden["cam0"]), torch.from_numpy(golden["K0"])
    mask = torch.from_numpy(golden["mask0"]).reshape(-1)
    origins, dirs = ray_store.pixel_rays(cam, K, mask, H=H, W=W)
    assert origins.shape == dirs.shape == (int(mask.sum()), 3) and dirs.dtype == torch.float32
    assert torch.equal(origins[0], cam[:, 3])
    n0 = golden["hit0_vids"].shape[0]
    assert torch.equal(dirs[torch.from_numpy(golden["hit0_hit_ray_idxs"])], torch.from_numpy(golden["loaded_unit_ray_dirs"][:n0]))
    # explicit pixel coordinates (the undistorted-pixels entry) reproduce the integer grid
    pix = torch.nonzero(mask).reshape(-1)
    _, d2 = ray_store.pixel_rays(cam, K, mask, H=H, W=W, pixel_coords=torch.stack([pix % W, pix // W], -1).float())
    assert torch.equal(d2, dirs)


def test_loader_matches_reference_load_preprocessed_data(golden, tmp_path):
    for stem in STEMS:
        (tmp_path / (stem + ".npy")).write_bytes(golden["file_" + stem].tobytes())
    names = ("vertex_idxs_of_hit_faces", "barycentric_coords", "expected_rgbs", "unit_ray_dirs", "face_idxs")
    wide = ray_store.load_preprocessed_rays(str(tmp_path), index_dtype=torch.int64)
    assert tuple(wide.keys()) == names                              # key order of dataset.py:12-33
    for n in names:
        ref = torch.from_numpy(golden["loaded_" + n])
        assert wide[n].dtype == ref.dtype and torch.equal(wide[n], ref), n
    for kw in (dict(), dict(mmap=True)):
        narrow = ray_store.load_preprocessed_rays(str(tmp_path), **kw)
        assert narrow["vertex_idxs_of_hit_faces"].dtype == torch.int32 and narrow["face_idxs"].dtype == torch.int32
        for n in names:
            assert torch.equal(narrow[n].to(wide[n].dtype), wide[n]), n
    # without the view files the reference returns three keys only
    os.remove(tmp_path / "face_idxs.npy")
    assert tuple(ray_store.load_preprocessed_rays(str(tmp_path)).keys()) == names[:3]


def test_writer_edge_cases(tmp_path):
    w = ray_store.RayCacheWriter(str(tmp_path / "empty"))
    w.write_to_disk()
    w.write_to_disk()                                               # idempotent
    d = ray_store.load_preprocessed_rays(str(tmp_path / "empty"))
    assert d["vertex_idxs_of_hit_faces"].shape == (0, 3) and d["face_idxs"].shape == (0,)
    for stem in STEMS:                                              # an empty store is still what np.save writes
        ref = tmp_path / "ref.npy"
        np.save(ref, np.zeros((0,) + ray_store.FILES[stem][1], dtype=ray_store.FILES[stem][0]), allow_pickle=False)
        assert (tmp_path / "empty" / (stem + ".npy")).read_bytes() == ref.read_bytes()

    w = ray_store.RayCacheWriter(str(tmp_path / "ragged"))
    g = torch.Generator().manual_seed(0)
    parts = []
    for n in (0, 1, 7, 0, 1000):                                    # views without a single hit are legal
        p = (torch.randint(0, 2**31 - 1, (n, 3), generator=g), torch.rand(n, 3, generator=g, dtype=torch.float64),
             torch.rand(n, 3, generator=g), torch.randn(n, 3, generator=g), torch.randint(0, 2**31 - 1, (n,), generator=g))
        parts.append(p)
        assert w.append_hits(*p) == n
    with pytest.raises(OverflowError):
        w.append_hits(torch.full((1, 3), 2**31), torch.rand(1, 3), torch.rand(1, 3), torch.rand(1, 3), torch.zeros(1, dtype=torch.int64))
    with pytest.raises(ValueError):
        w.append_hits(torch.zeros(2, 3, dtype=torch.int64), torch.rand(1, 3), torch.rand(2, 3), torch.rand(2, 3),
                      torch.zeros(2, dtype=torch.int64))
    w.write_to_disk()
    d = ray_store.load_preprocessed_rays(str(tmp_path / "ragged"), index_dtype=torch.int64)
    assert torch.equal(d["vertex_idxs_of_hit_faces"], torch.cat([p[0] for p in parts]))
    assert torch.equal(d["barycentric_coords"], torch.cat([p[1] for p in parts]).float())
    assert torch.equal(d["face_idxs"], torch.cat([p[4] for p in parts]))
    with pytest.raises(RuntimeError):
        w.cache_single_view(None, None, torch.ones(2, 2, dtype=torch.bool), torch.zeros(2, 2, 3))


@pytest.mark.parametrize("stem", STEMS)
def test_reserved_header_matches_numpy_for_every_realistic_row_count(stem):
    # the appender patches a 128-byte header in place; np.save must emit exactly that length for the same shape / dtype
    import io
    dt, tail = ray_store.FILES[stem]
    for n in (0, 1, 9, 10, 12345, 10**6, 10**9, 2**31 - 1, 10**12, 10**15):
        hdr = ray_store._npy_header(dt, (n,) + tail)
        assert len(hdr) == ray_store._HEADER_RESERVE, (stem, n, len(hdr))
        buf = io.BytesIO()
        np.lib.format.write_array_header_1_0(buf, {"descr": np.lib.format.dtype_to_descr(np.dtype(dt)), "fortran_order": False,
                                                   "shape": (n,) + tail})
        assert buf.getvalue() == hdr
    # and a real file of a few thousand rows equals np.save byte for byte
    import tempfile
    rows = np.arange(3000 * max(1, int(np.prod(tail))), dtype=dt).reshape((3000,) + tail)
    with tempfile.TemporaryDirectory() as d:
        a = ray_store.NpyAppender(os.path.join(d, "a.npy"), dt, tail)
        for lo in range(0, 3000, 701):
            a.append(rows[lo:lo + 701])
        a.close()
        np.save(os.path.join(d, "b.npy"), rows, allow_pickle=False)
        assert open(os.path.join(d, "a.npy"), "rb").read() == open(os.path.join(d, "b.npy"), "rb").read()


def test_load_first_k_eigenfunctions_matches_reference_bit_for_bit(tmp_path):
    # golden tables from the live reference (tests/golden/make_efuncs_golden.py): selection, GPS / HKS embedding, rescales
    z = np.load(os.path.join(GOLDEN_DIR, "efuncs.npz"))
    ep, vp = str(tmp_path / "efuncs.npy"), str(tmp_path / "evals.npy")
    np.save(ep, z["efuncs"])
    np.save(vp, z["evals"])
    cases = {
        "first_k_standard": dict(k=24, rescale_strategy="standard"),
        "list_k_standard": dict(k=[0, 1, 2, 7, 8, 30, 39], rescale_strategy="standard"),
        "first_k_unscaled": dict(k=40, rescale_strategy="unscaled"),
        "first_k_one_norm": dict(k=16, rescale_strategy="one-norm"),
        "gps": dict(k=20, rescale_strategy="standard", embed_strategy="gps", eigenvalues_path=vp),
        "hks_standard": dict(k=20, rescale_strategy="standard", embed_strategy="hks", eigenvalues_path=vp),
    }
    for name, kw in cases.items():
        k = kw.pop("k")
        got = ray_store.load_first_k_eigenfunctions(ep, k, **kw)
        want = torch.from_numpy(z["out_" + name]).to(torch.float32)          # the GPS branch of the reference returns float64
        assert got.dtype == torch.float32 and torch.equal(got, want), name
    with pytest.raises(RuntimeError):
        ray_store.load_first_k_eigenfunctions(ep, 4, rescale_strategy="bogus")
    with pytest.raises(ValueError):
        ray_store.load_first_k_eigenfunctions(ep, 4, embed_strategy="hks")        # no eigenvalues
    with pytest.raises(ValueError):
        ray_store.load_first_k_eigenfunctions(ep, 41)
