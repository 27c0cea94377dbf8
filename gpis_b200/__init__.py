"""gpis_b200: B200-native GPIS active-set selection + posterior prediction.

The compute path is hand-written sm_100a CUDA behind a C ABI (include/gpis_b200.h,
gpis_b200/libgpis_b200.so); this package is the Python host side that mirrors the
reference's operator interface.  There is no CPU fallback."""
from . import _lib
from .api import (GpModel, create_gpis, create_sparse_gpis, default_hyp, evaluate_errors, gp_cov,
                  gp_mean, predict_2d_grid, predict_grid, predict_points, sample_shapes, select_active_set,
                  select_active_set_straddle)
from .engine import GpisEngine
from .selector import GaussianProcessHyperparams, GpuActiveSetSelector, PredictionError
from .tsdf import auto_tsdf, build_tsdf

__all__ = ["GpisEngine", "GpuActiveSetSelector", "GaussianProcessHyperparams", "PredictionError", "GpModel", "create_gpis", "create_sparse_gpis", "default_hyp", "evaluate_errors",
           "gp_cov", "gp_mean", "predict_2d_grid", "predict_grid", "predict_points", "select_active_set",
           "select_active_set_straddle", "sample_shapes", "auto_tsdf", "build_tsdf"]
