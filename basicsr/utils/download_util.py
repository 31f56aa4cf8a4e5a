"""Cached weight download (reference: basicsr/utils/download_util.py:70-99).  Same contract: returns the
path under <torch hub dir>/checkpoints (or model_dir), downloading only when the file is absent."""
from __future__ import annotations

import os
from urllib.parse import urlparse


def load_file_from_url(url, model_dir=None, progress=True, file_name=None):
    import torch.hub as hub
    if model_dir is None:
        model_dir = os.path.join(hub.get_dir(), "checkpoints")
    os.makedirs(model_dir, exist_ok=True)
    name = file_name if file_name is not None else os.path.basename(urlparse(url).path)
    target = os.path.abspath(os.path.join(model_dir, name))
    if not os.path.exists(target):
        print(f'Downloading: "{url}" to {target}\n')
        hub.download_url_to_file(url, target, hash_prefix=None, progress=progress)
    return target
