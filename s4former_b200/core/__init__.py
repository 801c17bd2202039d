from .evaluation import (eval_metrics, intersect_and_union, mean_iou, total_area_to_metrics,
                         total_intersect_and_union)

__all__ = ['eval_metrics', 'intersect_and_union', 'mean_iou', 'total_area_to_metrics', 'total_intersect_and_union']
