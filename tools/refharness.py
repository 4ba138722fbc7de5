"""Development-only harness: import the upstream reference (drocheam/optrace) from
/root/reference with the three stubs SURVEY.md §8c describes.

Only the golden-vector generator (tools/gen_golden.py) and ad-hoc validation scripts use
this module. Nothing under tests/, bench.py or the product package may import it at GPU
run time: /root/reference does not exist on the GPU box.
"""
import sys
import types

REF_ROOT = "/root/reference"


def import_reference():
    """Returns the imported `optrace` module of the reference (with GUI/chardet stubbed)."""
    if "optrace" in sys.modules:
        return sys.modules["optrace"]

    # stub 1: traits.etsconfig.api (optrace/__init__.py:14-15 only sets ETSConfig.toolkit)
    traits = types.ModuleType("traits")
    ets = types.ModuleType("traits.etsconfig")
    api = types.ModuleType("traits.etsconfig.api")

    class ETSConfig:
        toolkit = None

    api.ETSConfig = ETSConfig
    traits.etsconfig = ets
    ets.api = api
    sys.modules.setdefault("traits", traits)
    sys.modules.setdefault("traits.etsconfig", ets)
    sys.modules.setdefault("traits.etsconfig.api", api)

    # stub 2: chardet (optrace/tracer/load.py:3, only used by the .zmx/.agf loaders)
    chardet = types.ModuleType("chardet")

    class EncodingEra:
        MODERN_WEB = 0

    def detect(b, **kw):
        if b[:2] in (b"\xff\xfe", b"\xfe\xff"):
            return {"encoding": "utf-16"}
        return {"encoding": "utf-8"}

    chardet.EncodingEra = EncodingEra
    chardet.detect = detect
    sys.modules.setdefault("chardet", chardet)

    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import optrace  # noqa

    # stub 3: optrace.gui (tests/tracing_geometry.py imports TraceGUI)
    gui = types.ModuleType("optrace.gui")

    class TraceGUI:
        def __init__(self, *a, **k):
            pass

        def run(self, *a, **k):
            pass

    gui.TraceGUI = TraceGUI
    sys.modules.setdefault("optrace.gui", gui)
    return optrace
