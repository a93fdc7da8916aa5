from tt_sketch.drm.sparse_gaussian_drm import SparseGaussianDRM
from tt_sketch.drm.sparse_sign_drm import SparseSignDRM
from tt_sketch.drm.tensor_train_drm import TensorTrainDRM

# DenseGaussianDRM of the reference is outside the accelerated path (not named by the north
# star); see DESIGN.md section 7.
ALL_DRM = (SparseGaussianDRM, TensorTrainDRM, SparseSignDRM)
