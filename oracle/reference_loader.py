"""TEST INFRASTRUCTURE ONLY (oracle).

Imports the UNMODIFIED reference package from /root/reference in THIS container
(the path does not exist on the GPU box) so that ``tests/golden/make_golden.py``
can (1) pin the oracle restatement against the real reference modules and
(2) dump golden vectors.  Nothing on the product path, in the ``-m gpu`` tests,
in ``smoke()`` or in ``bench.py`` imports this file.

Stub recipe:
  * ``accelerate`` / ``ema_pytorch`` are absent and only used by the trainers ->
    empty stand-ins;
  * ``vector_quantize_pytorch`` is absent -> ``oracle.lfq`` (restated LFQ);
  * T5 needs network weights -> ``get_encoded_dim`` patched to 768 and
    ``Phenaki.encode_texts`` replaced per instance by a synthetic embedder.
"""
import importlib.machinery
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "phenaki_pytorch"))


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    mod.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
    for k, v in attrs.items():
        setattr(mod, k, v)
    sys.modules[name] = mod
    return mod


def load_reference():
    """Returns the imported ``phenaki_pytorch`` reference package."""
    if "phenaki_pytorch" in sys.modules and getattr(sys.modules["phenaki_pytorch"], "_oracle_loaded", False):
        return sys.modules["phenaki_pytorch"]
    assert reference_available(), "reference tree not present (expected only in the build container)"
    import transformers  # noqa: F401  (must be imported before the accelerate stub exists)

    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.dirname(here))
    from oracle import lfq

    if "accelerate" not in sys.modules:
        _stub("accelerate", Accelerator=object, DistributedType=object,
              DistributedDataParallelKwargs=object)
    if "ema_pytorch" not in sys.modules:
        _stub("ema_pytorch", EMA=object)
    _stub("vector_quantize_pytorch", LFQ=lfq.LFQ, VectorQuantize=lfq.VectorQuantize)

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import phenaki_pytorch  # the reference
    import phenaki_pytorch.t5 as t5
    import phenaki_pytorch.phenaki_pytorch as pp

    t5.get_encoded_dim = lambda name: 768
    pp.get_encoded_dim = lambda name: 768
    phenaki_pytorch._oracle_loaded = True
    return phenaki_pytorch
