"""cardiax/plot.py is visualisation only (matplotlib) and outside the accelerated path.
This module keeps the names ``solve.forward(plot_while=True)`` and ``deepx.generate.sequence`` call -- ``plot_state``,
``plot_stimuli``, ``plot_diffusivity`` -- so that those branches degrade gracefully: with matplotlib installed they draw
a plain ``imshow`` panel per array, without it they say so once and return None.
"""
import warnings

import numpy as np

_warned = False


def _plt():
    global _warned
    try:
        import matplotlib.pyplot as plt
        return plt
    except Exception:  # matplotlib is absent from this image
        if not _warned:
            warnings.warn("cardiax.plot: matplotlib is not installed; plotting calls are skipped")
            _warned = True
        return None


def _np(a):
    return a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)


def _panels(arrays, titles, **kwargs):
    plt = _plt()
    if plt is None:
        return None
    fig, ax = plt.subplots(1, len(arrays), figsize=kwargs.pop("figsize", (5 * len(arrays), 5)))
    ax = np.atleast_1d(ax)
    for a, arr, title in zip(ax, arrays, titles):
        im = a.imshow(_np(arr), **kwargs)
        a.set_title(title)
        fig.colorbar(im, ax=a)
    return fig, ax


def plot_state(state, diffusivity=None, **kwargs):
    """cardiax/plot.py:43-89 -- v, w, u side by side."""
    return _panels(list(state), ["v", "w", "u"], **kwargs)


def plot_stimuli(stimuli, **kwargs):
    """cardiax/plot.py -- one panel per stimulus field."""
    if not len(stimuli):
        return None
    return _panels([s.field for s in stimuli], ["stimulus %d" % i for i in range(len(stimuli))], **kwargs)


def plot_diffusivity(diffusivity, **kwargs):
    return _panels([diffusivity], ["diffusivity"], **kwargs)
