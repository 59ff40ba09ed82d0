"""Public names of the package (mirrors what ``tdgl/__init__.py`` exports for the path)."""
from .mesh import EdgeMesh, Mesh, make_film_mesh

__all__ = ["Mesh", "EdgeMesh", "make_film_mesh"]
