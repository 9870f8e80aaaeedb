"""
`pmesh.window.methods` -- name -> resampler object with `.support`
(used at source/mesh/catalog.py:136,194,271-273; base/catalog.py:834,846).

Only windows with a CUDA scatter kernel can paint ('nnb'/'nearest', 'cic', 'tsc', 'pcs'); the
wavelet names of pmesh are kept as *names* so that `to_mesh(resampler='db6')` is accepted and
`compensated=True` raises the reference's "compensation ... is not defined" error
(source/mesh/tests/test_catalogmesh.py:133-145), but painting with them raises NotImplementedError.
"""


class ResampleWindow(object):
    def __init__(self, name, support, code):
        self.name = name
        self.support = support
        self.code = code          # NBK_WINDOW_* of include/nbk_b200.h, or None when there is no kernel

    def __repr__(self):
        return "ResampleWindow(%s, support=%d)" % (self.name, self.support)


methods = dict(
    nearest=ResampleWindow("nearest", 1, 1), nnb=ResampleWindow("nnb", 1, 1),
    linear=ResampleWindow("linear", 2, 2), cic=ResampleWindow("cic", 2, 2),
    quadratic=ResampleWindow("quadratic", 3, 3), tsc=ResampleWindow("tsc", 3, 3),
    cubic=ResampleWindow("cubic", 4, 4), pcs=ResampleWindow("pcs", 4, 4),
)
for _n, _s in [("lanczos2", 4), ("lanczos3", 6), ("db6", 6), ("db12", 12), ("db20", 20),
               ("sym6", 6), ("sym12", 12), ("sym20", 20)]:
    methods[_n] = ResampleWindow(_n, _s, None)


def FindResampler(window):
    if isinstance(window, ResampleWindow):
        return window
    if window not in methods:
        raise ValueError("unknown resampler %s" % str(window))
    return methods[window]
