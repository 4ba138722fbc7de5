"""Global options and the warning channel of the drop-in API.

Mirrors optrace/global_options.py:8-97 and optrace/warnings.py:13-32 of the reference.
`multithreading` is accepted for API compatibility but is a no-op here: all per-ray work runs
on the GPU (SURVEY.md §8b "Threading").
"""
from contextlib import contextmanager
import warnings as _warnings


class OptraceWarning(UserWarning):
    """Warning category used by the engine (name kept from the reference)."""


class _GlobalOptions:
    def __init__(self):
        self.__dict__["multithreading"] = True        # no-op, GPU engine
        self.__dict__["show_progress_bar"] = False    # no progress bars in this engine
        self.__dict__["show_warnings"] = True
        self.__dict__["wavelength_range"] = [380., 780.]
        self.__dict__["spectral_colormap"] = None
        self.__dict__["ui_dark_mode"] = True
        self.__dict__["plot_dark_mode"] = True

    def __setattr__(self, key, val):
        if key not in self.__dict__:
            raise AttributeError(f"Unknown option {key}.")
        if key in ("multithreading", "show_progress_bar", "show_warnings", "ui_dark_mode", "plot_dark_mode"):
            if not isinstance(val, bool):
                raise TypeError(f"Property '{key}' needs to be of type bool, but is {type(val)}.")
        if key == "wavelength_range":
            if not isinstance(val, (list, tuple)):
                raise TypeError(f"Property '{key}' needs to be of types list or tuple, but is {type(val)}.")
            if len(val) != 2:
                raise ValueError(f"{key} must have two elements.")
            if val[0] > 380.:
                raise ValueError(f"Property '{key}' needs to be below or equal to 380, but is {val[0]}.")
            if val[1] < 780.:
                raise ValueError(f"Property '{key}' needs to be above or equal to 780, but is {val[1]}.")
        self.__dict__[key] = val

    @contextmanager
    def no_warnings(self):
        old = self.show_warnings
        self.show_warnings = False
        try:
            yield
        finally:
            self.show_warnings = old

    @contextmanager
    def no_progress_bar(self):
        old = self.show_progress_bar
        self.show_progress_bar = False
        try:
            yield
        finally:
            self.show_progress_bar = old


global_options = _GlobalOptions()


def warning(text: str) -> None:
    """Emit an OptraceWarning unless warnings are switched off."""
    if global_options.show_warnings:
        _warnings.warn(text, OptraceWarning, stacklevel=2)
