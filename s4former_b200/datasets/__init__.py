from .pipeline import BranchPipeline, draw_pmd_params

__all__ = ['BranchPipeline', 'draw_pmd_params']
