"""Model descriptor shared by the host layer: which plugin objects a GP was built from
(reference: GP(D, covariance, mean, noise), gaussian_process.py:43-62)."""
from dataclasses import dataclass

COV_SE, COV_MATERN, COV_RQ = 0, 1, 2
MEAN_ZERO, MEAN_CONST, MEAN_NEGQUAD = 0, 1, 2


@dataclass(frozen=True)
class ModelSpec:
    D: int
    cov_kind: int = COV_SE
    degree: int = 0
    ard: bool = True
    mean_kind: int = MEAN_ZERO
    noise_params: tuple = (1, 0, 0)

    @property
    def cov_n(self):
        if not self.ard:
            return 3 if self.cov_kind == COV_RQ else 2
        return self.D + (2 if self.cov_kind == COV_RQ else 1)

    @property
    def noise_n(self):
        p = self.noise_params
        return int(p[0] == 1) + int(p[1] == 2) + 2 * int(p[2] == 1)

    @property
    def mean_n(self):
        return (0, 1, 1 + 2 * self.D)[self.mean_kind]

    @property
    def hyp_n(self):
        return self.cov_n + self.noise_n + self.mean_n
