from .array import ArrayCatalog
from .uniform import UniformCatalog, RandomCatalog

__all__ = ['ArrayCatalog', 'UniformCatalog', 'RandomCatalog']
