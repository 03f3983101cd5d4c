"""Hot-path configuration values, same environment names as the reference (``config.py:18-46``).

Only the fields the similarity-search path reads are kept: the k bounds (``config.py:28-30``), the
index folder name (``config.py:38``), the image extensions ``create_index`` globs (``config.py:39``), the
CLIP model name that fixes ``d`` (``config.py:25``) and ``BATCH_SIZE`` (``config.py:33`` -- dead in the
reference, used here for batched indexing).  Server, thumbnail and comment settings are out of scope.
"""
import os


class Config:
    CLIP_MODEL = os.getenv("EVOSSEARCH_CLIP_MODEL", "ViT-B/32")
    MIN_RESULTS = int(os.getenv("EVOSSEARCH_MIN_RESULTS", "3"))
    MAX_RESULTS = int(os.getenv("EVOSSEARCH_MAX_RESULTS", "48"))
    DEFAULT_RESULTS = int(os.getenv("EVOSSEARCH_DEFAULT_RESULTS", "12"))
    BATCH_SIZE = int(os.getenv("EVOSSEARCH_BATCH_SIZE", "32"))
    INDEX_FOLDER_NAME = os.getenv("EVOSSEARCH_INDEX_FOLDER", ".clip_index")
    SUPPORTED_EXTENSIONS = {".jpg", ".jpeg", ".png", ".bmp", ".webp"}
    # embedding width per CLIP model (oldapp.py:1088-1092)
    CLIP_DIMS = {"ViT-B/32": 512, "ViT-B/16": 512, "ViT-L/14": 768}


config = Config()
