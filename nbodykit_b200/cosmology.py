"""
A self-contained stand-in for the one thing the FFTPower path needs from `nbodykit.cosmology`: a
callable linear power spectrum P(k) to feed the mock generator (the reference gets it from CLASS via
classylss, which is not part of this path -- SURVEY.md §2.1 "cosmology: OUT OF SCOPE").

`NoWiggleEHPower` is the Eisenstein & Hu (1998) zero-baryon-oscillation fitting formula (their
eqs. 26, 28-31), normalised to sigma8 with a top-hat window; the same shape the reference offers as
`LinearPower(cosmo, z, transfer='NoWiggleEisensteinHu')` (cosmology/power/transfers.py:184-255).
"""
import numpy


class NoWiggleEHPower(object):
    def __init__(self, h=0.6774, Omega0_m=0.3089, Omega0_b=0.0486, n_s=0.9667, sigma8=0.8159, Tcmb=2.7255,
                 redshift=0.0, growth=1.0):
        self.h, self.Om, self.Ob, self.n_s, self.sigma8 = h, Omega0_m, Omega0_b, n_s, sigma8
        self.redshift = redshift
        self.growth = growth          # D(z)/D(0), supplied by the caller (no background solver here)
        self.theta = Tcmb / 2.7
        omh2 = self.Om * h * h
        obh2 = self.Ob * h * h
        fb = self.Ob / self.Om
        # sound horizon (EH98 eq. 26) and alpha_gamma (eq. 31)
        self.s = 44.5 * numpy.log(9.83 / omh2) / numpy.sqrt(1 + 10 * obh2 ** 0.75)
        self.alpha = 1 - 0.328 * numpy.log(431 * omh2) * fb + 0.38 * numpy.log(22.3 * omh2) * fb * fb
        self._norm = 1.0
        self._norm = (sigma8 / self.sigma_r(8.0)) ** 2

    def transfer(self, k):
        """k in h/Mpc"""
        k = numpy.asarray(k, dtype='f8')
        kmpc = k * self.h                       # 1/Mpc
        omh2 = self.Om * self.h ** 2
        gamma_eff = omh2 * (self.alpha + (1 - self.alpha) / (1 + (0.43 * kmpc * self.s) ** 4))   # eq. 30
        q = kmpc * self.theta ** 2 / gamma_eff                                                       # eq. 28
        L0 = numpy.log(2 * numpy.e + 1.8 * q)
        C0 = 14.2 + 731.0 / (1 + 62.5 * q)
        return L0 / (L0 + C0 * q * q)                                                                 # eq. 29

    def __call__(self, k):
        k = numpy.asarray(k, dtype='f8')
        with numpy.errstate(invalid='ignore', divide='ignore'):
            p = self._norm * self.growth ** 2 * k ** self.n_s * self.transfer(k) ** 2
        return numpy.where(k > 0, p, 0.0)

    def sigma_r(self, r, kmin=1e-5, kmax=1e2, n=4096):
        """rms of the top-hat smoothed linear field at radius r [Mpc/h]"""
        lnk = numpy.linspace(numpy.log(kmin), numpy.log(kmax), n)
        k = numpy.exp(lnk)
        x = k * r
        w = 3 * (numpy.sin(x) - x * numpy.cos(x)) / x ** 3
        integrand = k ** 3 * self(k) * w * w / (2 * numpy.pi ** 2)
        return numpy.sqrt(numpy.trapezoid(integrand, lnk))


LinearPower = NoWiggleEHPower
