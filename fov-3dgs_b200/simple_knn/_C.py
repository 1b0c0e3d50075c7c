"""`simple_knn._C` stand-in: distCUDA2(points[P,3] CUDA float) -> [P] mean squared distance to the 3 nearest neighbours
(reference: fov3dgs/submodules/simple-knn/spatial.cu:15-26, simple_knn.cu:63-218)."""
from fovgs import ops as _ops


def distCUDA2(points):
    return _ops.knn_mean_dist2(points)
