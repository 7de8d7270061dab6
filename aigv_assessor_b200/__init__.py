"""aigv_assessor_b200 — B200-native (sm_100a) vision hot path of AIGV-Assessor behind the reference's
``extract_feature`` seam. See DESIGN.md / INTEGRATION.md."""
from .config import VisionPathConfig, internvit_300m, internvit_6b, tiny  # noqa: F401

__all__ = ["VisionPathConfig", "internvit_300m", "internvit_6b", "tiny", "VisionEngine", "LanguageEngine", "ClipScorer",
           "install", "refresh", "uninstall"]


def __getattr__(name):  # torch-dependent pieces are imported lazily so config/synth stay importable anywhere
    if name == "VisionEngine":
        from .engine import VisionEngine
        return VisionEngine
    if name == "LanguageEngine":
        from .llm import LanguageEngine
        return LanguageEngine
    if name == "ClipScorer":
        from .scoring import ClipScorer
        return ClipScorer
    if name in ("metrics", "sharding", "tiling", "llm", "llm_glue", "scoring", "sampling", "slowfast"):
        import importlib
        return importlib.import_module("." + name, __name__)
    if name in ("install", "refresh", "uninstall"):
        from . import plugin
        return getattr(plugin, name)
    raise AttributeError(name)
