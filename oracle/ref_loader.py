"""TEST / MEASUREMENT INFRASTRUCTURE — imports the unmodified reference (drocheam/optrace) from oracle/_ref/
(placed there by tools/vendor_reference.py) with the three import stubs SURVEY.md 8c describes:
    traits.etsconfig.api   optrace/__init__.py:13-14 only sets ETSConfig.toolkit
    chardet                optrace/tracer/load.py:3, used by the .zmx / .agf loaders to sniff the encoding
    optrace.gui, optrace.plots   Qt / matplotlib front ends, never imported on the tracer path
Only bench.py's reference arm / cpu_baseline leg and tests may import this module; the product never does."""
import pathlib
import sys
import types

REF_ROOT = pathlib.Path(__file__).resolve().parent / "_ref"


def available() -> bool:
    return (REF_ROOT / "optrace" / "tracer" / "raytracer.py").exists()


def load():
    """the reference's `optrace` module"""
    if "optrace" in sys.modules:
        return sys.modules["optrace"]
    if not available():
        raise ImportError(f"reference not vendored under {REF_ROOT} (run tools/vendor_reference.py where /root/reference exists)")
    traits, ets, api = types.ModuleType("traits"), types.ModuleType("traits.etsconfig"), types.ModuleType("traits.etsconfig.api")

    class ETSConfig:
        toolkit = None

    api.ETSConfig = ETSConfig
    traits.etsconfig, ets.api = ets, api
    for name, mod in (("traits", traits), ("traits.etsconfig", ets), ("traits.etsconfig.api", api)):
        sys.modules.setdefault(name, mod)

    chardet = types.ModuleType("chardet")

    class EncodingEra:
        MODERN_WEB = 0

    def detect(b, **kw):
        return {"encoding": "utf-16" if b[:2] in (b"\xff\xfe", b"\xfe\xff") else "utf-8"}

    chardet.EncodingEra, chardet.detect = EncodingEra, detect
    sys.modules.setdefault("chardet", chardet)

    if str(REF_ROOT) not in sys.path:
        sys.path.insert(0, str(REF_ROOT))
    import optrace
    for sub in ("gui", "plots"):
        m = types.ModuleType(f"optrace.{sub}")
        sys.modules.setdefault(f"optrace.{sub}", m)
    return optrace
