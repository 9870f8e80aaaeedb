from .catalog import CatalogMesh
from .field import FieldMesh
from .array import ArrayMesh
from .species import MultipleSpeciesCatalogMesh

__all__ = ['CatalogMesh', 'FieldMesh', 'ArrayMesh', 'MultipleSpeciesCatalogMesh']
