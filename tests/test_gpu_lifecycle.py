"""GPU: the application's index lifecycle and search entry points (oldapp.py:54-135, :1972-2157) on the
.clip_index on-disk layout, with a deterministic stub in place of the CLIP encoder (no weights here)."""
import os
import pickle
import zlib

import numpy as np
import pytest

import evo_ssearch_b200 as evs
import oracle
from oracle import faiss_io

pytestmark = pytest.mark.gpu


class StubEncoder:
    """Same three functions as oldapp.py:30-52; embeddings are unit-norm functions of the file name / text."""

    d = 512

    def _vec(self, key: str):
        v = oracle.synth_fill(1, self.d, seed=zlib.crc32(key.encode()))[0]
        return v  # already L2-normalised, 1-D like .cpu().numpy().flatten()

    def get_image_embedding(self, image_path):
        name = os.path.basename(str(image_path))
        if name.startswith("broken"):
            raise OSError("cannot identify image file")
        return self._vec("img:" + name)

    def get_image_embedding_from_pil(self, pil_image):
        return self._vec("img:" + pil_image)

    def get_text_embedding(self, text):
        return self._vec("img:" + text) if text.endswith(".jpg") else self._vec("txt:" + text)


def _make_folder(tmp_path, n=40):
    names = [f"img_{i:03d}.jpg" for i in range(n)] + ["a.png", "b.webp", "broken_1.jpg", "notes.txt", "upper.JPG"]
    for i, nm in enumerate(names):
        (tmp_path / nm).write_bytes(b"x" * (i + 1))
    (tmp_path / "sub").mkdir()
    (tmp_path / "sub" / "nested.jpg").write_bytes(b"y")  # glob is non-recursive (oldapp.py:65)
    return names


def test_create_save_load_search(tmp_path):
    enc = StubEncoder()
    _make_folder(tmp_path)
    index, paths, meta = evs.create_index(tmp_path, enc)
    assert index.ntotal == len(paths) == len(meta) == 42  # 40 jpg + png + webp; broken/txt/JPG/nested skipped
    assert all(set(m) == {"path", "mtime", "size"} for m in meta)
    evs.save_index(index, paths, meta, tmp_path)
    ci = tmp_path / ".clip_index"
    assert sorted(p.name for p in ci.iterdir()) == ["index.faiss", "metadata.pkl", "paths.pkl"]
    # on-disk layout: faiss flat file with rows in paths order, plain pickles beside it
    d, n, metric, xb = faiss_io.read_index_flat(str(ci / "index.faiss"))
    assert (d, n, metric) == (512, 42, 0)
    want = np.stack([enc.get_image_embedding(p) for p in paths]).astype("float32")
    assert np.array_equal(xb, want)
    assert pickle.load(open(ci / "paths.pkl", "rb")) == paths

    evs.evict_index()  # cold load from disk
    index2, paths2, meta2 = evs.load_index(tmp_path)
    assert index2 is not None and index2.ntotal == 42 and paths2 == paths and meta2 == meta
    index3, _, _ = evs.load_index(tmp_path)
    assert index3 is index2  # resident: no re-read per request (the reference re-reads, oldapp.py:1993)

    # text search entry point: limit clamp, k = min(limit, N), descending similarity
    res = evs.search_text(tmp_path, "img_007.jpg", enc, limit=6)
    assert len(res) == 6 and res[0]["filename"] == "img_007.jpg" and abs(res[0]["similarity"] - 1.0) < 1e-6
    sims = [r["similarity"] for r in res]
    assert sims == sorted(sims, reverse=True)
    q = enc.get_text_embedding("img_007.jpg").reshape(1, -1)
    Dr, Ir = oracle.canon_search(q, want, 6)
    assert [r["path"] for r in res] == [paths[i] for i in Ir[0]]
    assert np.array_equal(np.array(sims, np.float32), Dr[0])
    assert len(evs.search_text(tmp_path, "a cat", enc, limit=1000)) == 12  # out of range -> DEFAULT_RESULTS
    assert len(evs.search_text(tmp_path, "a cat", enc, limit=48)) == 42  # k = min(limit, len(paths))
    by_time = evs.search_text(tmp_path, "a cat", enc, limit=12, sort_by="time")
    mt = [r["metadata"]["mtime"] for r in by_time]
    assert mt == sorted(mt, reverse=True)
    # image search entry point: by path and by "uploaded image"
    res_i = evs.search_image(tmp_path, tmp_path / "a.png", enc, limit=3)
    assert res_i[0]["filename"] == "a.png"
    assert evs.search_image(tmp_path, "b.webp", enc, limit=3)[0]["filename"] == "b.webp"
    assert evs.search_text(tmp_path / "sub", "x", enc) is None  # folder not indexed

    # re-index after a change: the resident copy is replaced when the file changes
    (tmp_path / "new.jpg").write_bytes(b"z")
    index4, paths4, meta4 = evs.create_index(tmp_path, enc)
    evs.save_index(index4, paths4, meta4, tmp_path)
    got, p5, _ = evs.load_index(tmp_path)
    assert got.ntotal == 43 and len(p5) == 43


def test_empty_folder_and_missing_metadata(tmp_path):
    enc = StubEncoder()
    assert evs.create_index(tmp_path, enc) == (None, None, None)  # no images (oldapp.py:82-83)
    (tmp_path / "one.jpg").write_bytes(b"1")
    index, paths, meta = evs.create_index(tmp_path, enc)
    evs.save_index(index, paths, meta, tmp_path)
    os.remove(tmp_path / ".clip_index" / "metadata.pkl")  # backwards compatible (oldapp.py:124-131)
    evs.evict_index()
    i2, p2, m2 = evs.load_index(tmp_path)
    assert i2.ntotal == 1 and p2 == paths and m2 is None
    res = evs.search_text(tmp_path, "anything", enc, limit=12)
    assert len(res) == 1 and res[0]["metadata"] == {}


class BatchEncoder(StubEncoder):
    """Adds the optional batched protocol: ``encode_images(paths)`` -> raw (un-normalised) features on the GPU,
    like one ``model.encode_image`` call on a stacked batch."""

    def __init__(self, as_tensor=True):
        self.as_tensor = as_tensor
        self.batches = []
        self.single_calls = 0

    def raw(self, name):
        return oracle.synth_fill(1, self.d, seed=zlib.crc32(("raw:" + name).encode()), normalize=False)[0] * 3.5

    def get_image_embedding(self, image_path):
        self.single_calls += 1
        name = os.path.basename(str(image_path))
        if name.startswith("broken"):
            raise OSError("cannot identify image file")
        return oracle.l2_normalize(self.raw(name)[None, :])[0]

    def encode_images(self, paths):
        names = [os.path.basename(p) for p in paths]
        self.batches.append(len(names))
        if any(n.startswith("broken") for n in names):
            raise OSError("cannot identify image file")  # the whole batch fails, like a stacked forward would
        raw = np.stack([self.raw(n) for n in names])
        if self.as_tensor:
            import torch
            return torch.from_numpy(raw).cuda()
        return raw


@pytest.mark.parametrize("as_tensor", [True, False])
def test_batched_indexing_equals_per_image(tmp_path, as_tensor):
    _make_folder(tmp_path)
    enc = BatchEncoder(as_tensor)
    index, paths, meta = evs.create_index(tmp_path, enc, batch_size=7)
    assert index.ntotal == len(paths) == len(meta) == 42
    assert max(enc.batches) == 7 and sum(enc.batches) >= 42
    assert 0 < enc.single_calls <= 7  # only the batch holding the unreadable file was retried one by one
    # rows = our normalise kernel applied to the raw batch = the oracle's normaliser, bit for bit
    want = np.stack([oracle.l2_normalize(enc.raw(os.path.basename(p))[None, :])[0] for p in paths])
    assert np.array_equal(index.reconstruct_n(0, 42), want)
    # and the same index as the reference's one-image-at-a-time protocol builds
    ref_index, ref_paths, _ = evs.create_index(tmp_path, StubNoBatch(enc), batch_size=1)
    assert ref_paths == paths and np.array_equal(ref_index.reconstruct_n(0, 42), want)


class StubNoBatch:
    def __init__(self, inner):
        self.inner = inner

    def get_image_embedding(self, p):
        return self.inner.get_image_embedding(p)


def test_incremental_reindex_equals_fresh_index(tmp_path):
    enc = BatchEncoder()
    _make_folder(tmp_path)
    index, paths, meta = evs.create_index(tmp_path, enc)
    evs.save_index(index, paths, meta, tmp_path)
    # nothing changed: every row is kept, nothing is embedded
    enc.batches.clear()
    enc.single_calls = 0
    i1, p1, m1, st = evs.update_index(tmp_path, enc)
    assert st == {"kept": 42, "embedded": 0, "removed": 0, "failed": 1} and enc.batches == [1]  # broken_1.jpg is re-tried
    assert p1 == paths and m1 == meta and np.array_equal(i1.reconstruct_n(0, 42), index.reconstruct_n(0, 42))
    # delete two, add three, change one (size differs -> re-embedded)
    os.remove(tmp_path / "img_003.jpg")
    os.remove(tmp_path / "a.png")
    for nm in ("new_1.jpg", "new_2.png", "new_3.webp"):
        (tmp_path / nm).write_bytes(b"n" * 7)
    (tmp_path / "img_010.jpg").write_bytes(b"changed-content")
    enc.batches.clear()
    enc.single_calls = 0
    i2, p2, m2, st = evs.update_index(tmp_path, enc, batch_size=16)
    assert st["kept"] == 39 and st["embedded"] == 4 and st["removed"] == 2 and st["failed"] == 1
    assert sum(enc.batches) + enc.single_calls <= 5 + 5  # 5 files to embed (one unreadable), nothing else touched
    fresh, pf, mf = evs.create_index(tmp_path, BatchEncoder())
    assert p2 == pf and m2 == mf
    assert np.array_equal(i2.reconstruct_n(0, i2.ntotal), fresh.reconstruct_n(0, fresh.ntotal))
    evs.save_index(i2, p2, m2, tmp_path)
    res = evs.search_text(tmp_path, "x", enc, limit=48)
    assert len(res) == 43
    # no index yet -> behaves like create_index
    other = tmp_path / "other"
    other.mkdir()
    (other / "one.jpg").write_bytes(b"1")
    i3, p3, m3, st3 = evs.update_index(other, enc)
    assert i3.ntotal == 1 and st3["embedded"] == 1 and st3["kept"] == 0
    # row gather argument checks
    with pytest.raises(evs.EvsError):
        i3.add_rows_from(i2, [i2.ntotal])
    with pytest.raises(evs.EvsError):
        i3.add_rows_from(i3, [0])
