"""On-disk formats: load_fold against a line-by-line restatement of the
reference's fold-loop loading code (evel_her2st.py:126-172) on temporary files."""
import os

import numpy as np
import pytest

from mclstexp_b200 import io as mio
from mclstexp_b200.distributed import shard_bounds


def _make_files(tmp_path, sizes, genes, dim=256, seed=0, expr_dtype=np.float64):
    rng = np.random.default_rng(seed)
    emb_dir = os.path.join(tmp_path, "embeddings_0")
    n = sum(sizes)
    img_all = rng.standard_normal((n, dim)).astype(np.float32)
    spot_all = rng.standard_normal((n, dim)).astype(np.float32)
    mio.save_fold_embeddings(emb_dir, img_all, spot_all, sizes)
    paths = []
    for i, s in enumerate(sizes):
        d = os.path.join(tmp_path, f"slide{i}")
        os.makedirs(d)
        p = os.path.join(d, "preprocessed_matrix.npy")
        np.save(p, rng.random((genes, s)).astype(expr_dtype))          # [G, n_i]  (hvg_her2st.py:123)
        paths.append(p)
    return emb_dir, paths, img_all, spot_all


def _reference_fold(emb_dir, paths, fold):
    """evel_her2st.py:136-172 restated (paths instead of the hard-coded directories)."""
    n_slides = len(paths)
    spot_expressions = [np.load(p) for p in paths]
    spot_embeddings = [np.load(os.path.join(emb_dir, f"spot_embeddings_{i + 1}.npy")) for i in range(n_slides)]
    image_embeddings = np.load(os.path.join(emb_dir, f"img_embeddings_{fold + 1}.npy"))
    image_query = image_embeddings
    expression_gt = spot_expressions[fold]
    spot_embeddings = spot_embeddings[:fold] + spot_embeddings[fold + 1:]
    spot_expressions_rest = spot_expressions[:fold] + spot_expressions[fold + 1:]
    spot_key = np.concatenate(spot_embeddings, axis=1)
    expression_key = np.concatenate(spot_expressions_rest, axis=1)
    if image_query.shape[1] != 256:
        image_query = image_query.T
    if expression_gt.shape[0] != image_query.shape[0]:
        expression_gt = expression_gt.T
    if spot_key.shape[1] != 256:
        spot_key = spot_key.T
    if expression_key.shape[0] != spot_key.shape[0]:
        expression_key = expression_key.T
    return spot_key, expression_key, image_query, expression_gt


@pytest.mark.parametrize("fold", [0, 2, 4])
@pytest.mark.parametrize("mmap", [True, False])
def test_load_fold_matches_reference_loading(tmp_path, fold, mmap):
    sizes = [37, 120, 64, 301, 5]
    emb_dir, paths, _, _ = _make_files(str(tmp_path), sizes, genes=41, seed=fold)
    assert mio.slide_sizes(paths) == sizes
    ref = _reference_fold(emb_dir, paths, fold)
    got = mio.load_fold(emb_dir, paths, fold, mmap=mmap)
    for a, b in zip((got.spot_key, got.expression_key, got.image_query, got.expression_gt), ref):
        assert a.shape == b.shape
        np.testing.assert_array_equal(a, b)
    assert got.index_offset == 0 and got.n_total == sum(sizes) - sizes[fold]
    assert got.spot_key.dtype == np.float32 and got.expression_key.dtype == np.float64
    assert got.spot_key.flags.c_contiguous and got.image_query.flags.c_contiguous


@pytest.mark.parametrize("world", [2, 3, 8])
def test_load_fold_shards_concatenate_to_the_full_bank(tmp_path, world):
    sizes = [50, 7, 211, 96]
    emb_dir, paths, _, _ = _make_files(str(tmp_path), sizes, genes=17, seed=9, expr_dtype=np.float32)
    full = mio.load_fold(emb_dir, paths, 1)
    parts = [mio.load_fold(emb_dir, paths, 1, rank=r, world=world, expression_dtype=np.float32)
             for r in range(world)]
    bounds = shard_bounds(full.n_total, world)
    for r, part in enumerate(parts):
        assert (part.index_offset, part.index_offset + part.spot_key.shape[0]) == bounds[r]
        assert part.n_total == full.n_total
        np.testing.assert_array_equal(part.image_query, full.image_query)
    np.testing.assert_array_equal(np.concatenate([p.spot_key for p in parts]), full.spot_key)
    np.testing.assert_array_equal(np.concatenate([p.expression_key for p in parts]), full.expression_key)


def test_row_major_files_and_errors(tmp_path):
    sizes = [30, 40]
    emb_dir, paths, img_all, spot_all = _make_files(str(tmp_path), sizes, genes=9)
    # a slide stored [n, 256] instead of [256, n] is accepted
    np.save(os.path.join(emb_dir, "spot_embeddings_2.npy"), spot_all[30:70])
    got = mio.load_fold(emb_dir, paths, 0)
    np.testing.assert_array_equal(got.spot_key, spot_all[30:70])
    np.testing.assert_array_equal(got.image_query, img_all[:30])
    with pytest.raises(ValueError):
        mio.load_fold(emb_dir, paths, 2)
    with pytest.raises(ValueError):
        mio.load_fold(emb_dir, paths, 0, rank=2, world=2)
    np.save(os.path.join(emb_dir, "spot_embeddings_2.npy"), spot_all[30:69].T)      # one spot short
    with pytest.raises(ValueError):
        mio.load_fold(emb_dir, paths, 0)
    with pytest.raises(ValueError):
        mio.save_fold_embeddings(emb_dir, img_all, spot_all, [30, 41])


def test_writer_and_loader_match_reference_executed_digests(tmp_path):
    """tests/golden/io.json: SHA-256 digests of the four arrays the reference's OWN writer loop
    (evel_her2st.py:109-117) and fold-loop loading code (:145-172) produce, executed verbatim by
    oracle/make_golden.py on seeded inputs.  Our writer + loader must reproduce them byte for byte."""
    import hashlib
    import json
    from oracle import make_golden                            # only io_inputs(): the seeded inputs
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, "tests", "golden", "io.json")) as f:
        gold = json.load(f)
    img, spot, mats = make_golden.io_inputs()
    assert gold["sizes"] == make_golden.IO_SIZES and gold["genes"] == make_golden.IO_GENES
    paths = []
    for i, m in enumerate(mats):
        d = os.path.join(str(tmp_path), f"slide{i}")
        os.makedirs(d)
        np.save(os.path.join(d, "preprocessed_matrix.npy"), m)
        paths.append(os.path.join(d, "preprocessed_matrix.npy"))
    for fold in (0, 17, 31):
        emb_dir = os.path.join(str(tmp_path), f"embeddings_{fold}")
        mio.save_fold_embeddings(emb_dir, img, spot, gold["sizes"])
        fd = mio.load_fold(emb_dir, paths, fold)
        for key in ("spot_key", "expression_key", "image_query", "expression_gt"):
            a = np.ascontiguousarray(getattr(fd, key))
            ref = gold["arrays"][f"{fold}/{key}"]
            assert list(a.shape) == ref["shape"] and str(a.dtype) == ref["dtype"], (fold, key)
            assert hashlib.sha256(a.tobytes()).hexdigest() == ref["sha256"], (fold, key)
