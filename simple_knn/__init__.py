"""Drop-in import name for the second native dependency of the reference's Gaussian model:
`from simple_knn._C import distCUDA2` (gs-simp/scene/gaussian_model.py:20).  B200-native implementation in
multiview_inpaint_b200/csrc/knn.cu behind the C ABI gsr_knn3_mean_dist2 (include/gsrast_b200.h)."""
