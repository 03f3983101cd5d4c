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
