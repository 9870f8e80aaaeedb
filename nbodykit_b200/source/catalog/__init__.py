from .array import ArrayCatalog
from .uniform import UniformCatalog, RandomCatalog
from .lognormal import LogNormalCatalog

__all__ = ["ArrayCatalog", "UniformCatalog", "RandomCatalog", "LogNormalCatalog"]
