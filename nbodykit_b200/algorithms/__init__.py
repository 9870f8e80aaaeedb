from .fftpower import FFTPower, FFTBase, project_to_basis

__all__ = ['FFTPower', 'FFTBase', 'project_to_basis']
