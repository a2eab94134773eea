import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _native_built():
    # the oracle and the host-logic double are CPU code and are (re)built on demand; the CUDA library is
    # built by __graft_entry__.build() and travels to the GPU box as a prebuilt .so
    import backends
    backends.build()
    backends.build_hostdouble()
    yield


@pytest.fixture(scope="session")
def ab_comm():
    """(package, comm handle) for the PARPACK entry points on the GPU: ONE 1-rank NCCL process group and ONE library
    communicator for the whole session, shared by every test module that needs them (no re-initialisation of NCCL
    inside a process)."""
    import torch
    import torch.distributed as dist
    import arpack_ng_b200 as ab
    ab.lib()
    created = False
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", str(29600 + os.getpid() % 1500))
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
        created = True
    comm = ab.nccl_comm_from_torch_distributed()
    yield ab, comm
    ab.lib().ab200_comm_destroy(comm)
    if created:
        dist.destroy_process_group()
