"""Import the REAL reference files (attend.py, audiolm_pytorch.py, soundstream.py) from /root/reference.

TEST INFRASTRUCTURE, build container only: /root/reference does not exist on the GPU box, so nothing
under tests/ (-m gpu), smoke() or bench.py may call this.  It exists to (a) generate tests/golden/*
and (b) pin oracle/transformer.py and oracle/codec.py against the reference's own code.

Nine third-party packages the reference imports are absent offline.  They are replaced
by stub modules; the two that carry hot-path arithmetic get the restatements in oracle/third_party.py.
"""
from __future__ import annotations

import sys
import types
from pathlib import Path
from types import SimpleNamespace

from torch import nn

REFERENCE_ROOT = Path("/root/reference")


class _Absent(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        self.kwargs = k

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("this third-party module is not available offline")


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_loaded = None


def available() -> bool:
    return (REFERENCE_ROOT / "audiolm_pytorch" / "audiolm_pytorch.py").exists()


def load():
    """Returns SimpleNamespace(lm=<audiolm_pytorch.audiolm_pytorch>, ss=<...soundstream>, attend=<...attend>)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("/root/reference is not present (GPU box?) — goldens live in tests/golden/")
    from . import third_party as tp

    _stub("fairseq")
    _stub("encodec", EncodecModel=_Absent)
    _stub("encodec.utils", _linear_overlap_add=None)
    _stub("vector_quantize_pytorch", GroupedResidualVQ=tp.GroupedResidualVQ, GroupedResidualLFQ=_Absent,
          GroupedResidualFSQ=_Absent, ResidualVQ=tp.ResidualVQ)
    _stub("local_attention", LocalMHA=tp.LocalMHA)
    _stub("local_attention.transformer", FeedForward=tp.LocalFeedForward, DynamicPositionBias=_Absent)
    _stub("gateloop_transformer", SimpleGateLoopLayer=_Absent)
    _stub("hyper_connections",
          get_init_and_expand_reduce_stream_functions=tp.get_init_and_expand_reduce_stream_functions)
    pkg = types.ModuleType("audiolm_pytorch")
    pkg.__path__ = [str(REFERENCE_ROOT / "audiolm_pytorch")]  # package __init__ (-> trainer -> accelerate) never runs
    sys.modules["audiolm_pytorch"] = pkg

    import audiolm_pytorch.t5 as t5  # noqa: E402

    t5.T5_CONFIGS[t5.DEFAULT_T5_NAME] = dict(config=SimpleNamespace(d_model=768))  # no HF hub access in ctors
    import audiolm_pytorch.attend as ref_attend  # noqa: E402
    import audiolm_pytorch.audiolm_pytorch as ref_lm  # noqa: E402
    import audiolm_pytorch.soundstream as ref_ss  # noqa: E402

    _loaded = SimpleNamespace(lm=ref_lm, ss=ref_ss, attend=ref_attend)
    return _loaded
