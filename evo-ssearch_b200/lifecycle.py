"""Index lifecycle and search entry points of the application, kept by name.

Mirrors ``oldapp.py``::

    create_index(folder)                         oldapp.py:54-90
    save_index(index, paths, metadata, folder)   oldapp.py:92-106
    load_index(folder)                           oldapp.py:108-135
    search (text) handler body                   oldapp.py:1985-2045  -> search_text
    search_by_image handler body                 oldapp.py:2065-2152  -> search_image

with the same ``<folder>/.clip_index/{index.faiss, paths.pkl, metadata.pkl}`` files.  What changes is the
engine underneath (``IndexFlatIP`` on the GPU) and that ``load_index`` keeps indexes RESIDENT: the
reference re-reads the whole ``index.faiss`` on every request (oldapp.py:1993, :2084); here a loaded
index stays in HBM, keyed by the file's (path, mtime, size) and reloaded only when that changes.

The CLIP encoder is out of scope (north_star: "unchanged CLIP ViT encoder"); callers pass any object
with the reference's three embedding functions (``get_image_embedding``, ``get_image_embedding_from_pil``,
``get_text_embedding``; oldapp.py:30-52).  HTTP, thumbnails and JSON stay in the application.
"""
from __future__ import annotations

import os
import pickle
import threading
from pathlib import Path
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from .config import config
from .index import IndexFlatIP, read_index, write_index

_cache_lock = threading.Lock()
_cache: Dict[str, Tuple[Tuple[int, int], IndexFlatIP, list, Optional[list]]] = {}


def create_index(folder_path, encoder) -> Tuple[Optional[IndexFlatIP], Optional[List[str]], Optional[List[dict]]]:
    """Embed every supported image directly inside ``folder_path`` and build the index (oldapp.py:54-90).

    Same walk as the reference: one non-recursive glob per extension, per-image failures are printed
    and skipped, ``(None, None, None)`` when nothing could be embedded.  Embeddings are stacked and
    cast to float32 exactly as ``np.array(embeddings).astype('float32')`` (oldapp.py:86).
    """
    folder_path = Path(folder_path)
    image_paths: List[str] = []
    embeddings: List[np.ndarray] = []
    image_metadata: List[dict] = []
    for ext in config.SUPPORTED_EXTENSIONS:
        for img_path in folder_path.glob(f"*{ext}"):
            try:
                embedding = encoder.get_image_embedding(img_path)
                embeddings.append(embedding)
                image_paths.append(str(img_path))
                stat = img_path.stat()
                image_metadata.append({"path": str(img_path), "mtime": stat.st_mtime, "size": stat.st_size})
            except Exception as e:  # noqa: BLE001 - the reference prints and continues
                print(f"Error processing {img_path}: {e}")
    if not embeddings:
        return None, None, None
    embeddings_array = np.array(embeddings).astype("float32")
    index = IndexFlatIP(embeddings_array.shape[1])
    index.add(embeddings_array)
    return index, image_paths, image_metadata


def save_index(index: IndexFlatIP, image_paths, image_metadata, folder_path) -> None:
    """Write ``index.faiss``, ``paths.pkl``, ``metadata.pkl`` under ``<folder>/.clip_index`` (oldapp.py:92-106)."""
    index_path = Path(folder_path) / config.INDEX_FOLDER_NAME
    index_path.mkdir(exist_ok=True)
    write_index(index, str(index_path / "index.faiss"))
    with open(index_path / "paths.pkl", "wb") as f:
        pickle.dump(image_paths, f)
    with open(index_path / "metadata.pkl", "wb") as f:
        pickle.dump(image_metadata, f)
    # the index just written is the resident one for this folder
    key = str((index_path / "index.faiss").resolve())
    st = os.stat(key)
    with _cache_lock:
        _cache[key] = ((st.st_mtime_ns, st.st_size), index, list(image_paths), image_metadata)


def load_index(folder_path):
    """``(index, image_paths, image_metadata)`` or ``(None, None, None)`` (oldapp.py:108-135).

    Any failure -- missing folder, corrupt file, unreadable pickle -- yields the ``None`` triple, as the
    reference's blanket ``except`` does.  A hit in the resident cache costs one ``stat``.
    """
    index_path = Path(folder_path) / config.INDEX_FOLDER_NAME
    if not index_path.exists():
        return None, None, None
    try:
        fname = index_path / "index.faiss"
        key = str(fname.resolve())
        st = os.stat(key)
        sig = (st.st_mtime_ns, st.st_size)
        with _cache_lock:
            hit = _cache.get(key)
        if hit is not None and hit[0] == sig:
            return hit[1], hit[2], hit[3]
        index = read_index(str(fname))
        with open(index_path / "paths.pkl", "rb") as f:
            image_paths = pickle.load(f)
        image_metadata = None
        metadata_file = index_path / "metadata.pkl"
        if metadata_file.exists():
            try:
                with open(metadata_file, "rb") as f:
                    image_metadata = pickle.load(f)
            except Exception:  # noqa: BLE001 - backwards compatible, as the reference
                image_metadata = None
        with _cache_lock:
            _cache[key] = (sig, index, image_paths, image_metadata)
        return index, image_paths, image_metadata
    except Exception:  # noqa: BLE001 - the reference swallows everything here
        return None, None, None


def evict_index(folder_path=None) -> None:
    """Drop one folder's resident index (or all of them) and free its HBM."""
    with _cache_lock:
        if folder_path is None:
            _cache.clear()
        else:
            key = str((Path(folder_path) / config.INDEX_FOLDER_NAME / "index.faiss").resolve())
            _cache.pop(key, None)


def clamp_limit(limit: Any) -> int:
    """The handlers' limit rule (oldapp.py:1985-1990, :2065-2070): int in [MIN, MAX] else DEFAULT."""
    try:
        limit = int(limit)
        if limit < config.MIN_RESULTS or limit > config.MAX_RESULTS:
            limit = config.DEFAULT_RESULTS
    except (ValueError, TypeError):
        limit = config.DEFAULT_RESULTS
    return limit


def _collect(index, image_paths, image_metadata, embedding, limit, sort_by) -> List[dict]:
    """Search + the handlers' post-processing (oldapp.py:2002-2045) minus thumbnails."""
    k = min(limit, len(image_paths))
    if k == 0:
        return []
    similarities, indices = index.search(np.asarray(embedding).reshape(1, -1), k)
    results = []
    for idx, sim in zip(indices[0], similarities[0]):
        if idx >= 0 and idx < len(image_paths):
            img_path = image_paths[idx]
            metadata_info = {}
            if image_metadata and idx < len(image_metadata):
                meta = image_metadata[idx]
                metadata_info = {"mtime": meta.get("mtime", 0), "size": meta.get("size", 0)}
            results.append({"path": img_path, "filename": os.path.basename(img_path), "similarity": float(sim),
                            "metadata": metadata_info})
    if sort_by == "time" and image_metadata:
        results.sort(key=lambda x: x["metadata"].get("mtime", 0), reverse=True)
    return results


def search_text(folder, query: str, encoder, limit=10, sort_by: str = "similarity") -> Optional[List[dict]]:
    """Body of ``POST /search`` (oldapp.py:1972-2053).  ``None`` = folder not indexed."""
    limit = clamp_limit(limit)
    index, image_paths, image_metadata = load_index(folder)
    if index is None:
        return None
    return _collect(index, image_paths, image_metadata, encoder.get_text_embedding(query), limit, sort_by)


def search_image(folder, image, encoder, limit=12, sort_by: str = "similarity") -> Optional[List[dict]]:
    """Body of ``POST /search_by_image`` (oldapp.py:2055-2157).  ``image`` is a path or a PIL image."""
    limit = clamp_limit(limit)
    index, image_paths, image_metadata = load_index(folder)
    if index is None:
        return None
    if isinstance(image, (str, os.PathLike)):
        emb = encoder.get_image_embedding(image)
    else:
        emb = encoder.get_image_embedding_from_pil(image)
    return _collect(index, image_paths, image_metadata, emb, limit, sort_by)
