from .catalog import CatalogMesh
from .field import FieldMesh
from .array import ArrayMesh

__all__ = ['CatalogMesh', 'FieldMesh', 'ArrayMesh']
