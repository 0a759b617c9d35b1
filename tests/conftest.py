import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REFERENCE = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: imports the read-only reference checkout (this container only)")


def have_reference():
    return os.path.isdir(os.path.join(REFERENCE, "onnxocr"))


@pytest.fixture(scope="session")
def model_paths():
    from onnxocr_b200 import synth_graphs
    paths = synth_graphs.ensure_models()
    paths["cls"] = os.path.join(synth_graphs.ASSET_DIR, "ppocrv5_cls.onnx")
    paths["dict"] = os.path.join(synth_graphs.ASSET_DIR, "ppocrv5_dict.txt")
    return paths


@pytest.fixture(scope="session")
def lib_built():
    from onnxocr_b200 import build
    return build.build()


@pytest.fixture(scope="session")
def engine_fp32(lib_built, model_paths):
    from onnxocr_b200 import _lib
    e = _lib.Engine(0)
    e.load_onnx(_lib.DET, model_paths["det"], _lib.FP32)
    e.load_onnx(_lib.CLS, model_paths["cls"], _lib.FP32)
    e.load_onnx(_lib.REC, model_paths["rec"], _lib.FP32)
    yield e
    e.close()


@pytest.fixture(scope="session")
def engine_bf16(lib_built, model_paths):
    from onnxocr_b200 import _lib
    e = _lib.Engine(0)
    e.load_onnx(_lib.DET, model_paths["det"], _lib.BF16)
    e.load_onnx(_lib.CLS, model_paths["cls"], _lib.BF16)
    e.load_onnx(_lib.REC, model_paths["rec"], _lib.BF16)
    yield e
    e.close()


@pytest.fixture(scope="session")
def torch_graphs(model_paths):
    from oracle.torch_interp import TorchGraph
    cache = {}

    def get(kind):
        if kind not in cache:
            cache[kind] = TorchGraph(model_paths[kind])
        return cache[kind]
    return get


def rng(seed):
    return np.random.default_rng(seed)
