"""Helpers the inference entry point imports from `basicsr.utils` (inference_femasr.py:9)."""
import logging
import os

from .img_util import crop_border, imfrombytes, img2tensor, imwrite, tensor2img

__all__ = ["img2tensor", "tensor2img", "imwrite", "imfrombytes", "crop_border", "get_root_logger", "scandir"]


def get_root_logger(logger_name="basicsr", log_level=logging.INFO, log_file=None):
    logger = logging.getLogger(logger_name)
    if not logger.handlers:
        handler = logging.StreamHandler()
        handler.setFormatter(logging.Formatter("%(asctime)s %(levelname)s: %(message)s"))
        logger.addHandler(handler)
        logger.propagate = False
        logger.setLevel(log_level)
    if log_file is not None:
        fh = logging.FileHandler(log_file, "w")
        fh.setLevel(log_level)
        logger.addHandler(fh)
    return logger


def scandir(dir_path, suffix=None, recursive=False, full_path=False):
    for root, _dirs, files in os.walk(dir_path):
        for f in sorted(files):
            if not f.startswith(".") and (suffix is None or f.endswith(suffix)):
                p = os.path.join(root, f)
                yield p if full_path else os.path.relpath(p, dir_path)
        if not recursive:
            break
