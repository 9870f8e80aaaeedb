from .catalog import FKPCatalog, FKPWeightFromNbar
from .catalogmesh import FKPCatalogMesh
from .fkp import ConvolvedFFTPower

__all__ = ['FKPCatalog', 'FKPWeightFromNbar', 'FKPCatalogMesh', 'ConvolvedFFTPower']
