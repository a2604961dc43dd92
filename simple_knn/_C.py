"""Mirror of the simple-knn extension's pybind module: one function, `distCUDA2(points) -> Tensor[P]`
(reference call sites: gs-simp/scene/gaussian_model.py:134, :546, :623)."""
from multiview_inpaint_b200._C import dist_cuda2 as distCUDA2  # noqa: F401
