"""Values the prediction path shares with the reference
(volume_segmantics/utilities/config.py).  Only what prediction needs."""

# command-line argument names (arg_parsing.py)
MODEL_PTH_ARG, PREDICT_DATA_ARG, DATA_DIR_ARG = "model", "data", "data_dir"
TRAIN_DATA_ARG, LABEL_DATA_ARG = "data", "labels"

# accepted file suffixes
HDF5_SUFFIXES = {".h5", ".hdf5", ".nxs"}
TIFF_SUFFIXES = {".tiff", ".tif"}
PREDICT_DATA_EXT = HDF5_SUFFIXES | TIFF_SUFFIXES
TRAIN_DATA_EXT = LABEL_DATA_EXT = PREDICT_DATA_EXT
MODEL_DATA_EXT = {".pytorch", ".pth"}

# logging
LOGGING_FMT = "%(asctime)s - %(levelname)s - %(message)s"
LOGGING_DATE_FMT = "%d-%b-%y %H:%M:%S"
TQDM_BAR_FORMAT = "{l_bar}{bar: 30}{r_bar}{bar: -30b}"

# settings files
SETTINGS_DIR = "volseg-settings"
PREDICTION_SETTINGS_FN = "2d_model_predict_settings.yaml"
TRAIN_SETTINGS_FN = "2d_model_train_settings.yaml"

HDF5_COMPRESSION = "gzip"

# reference batch policy (get_batch_size); the B200 engine sizes its own batches
BIG_CUDA_THRESHOLD = 8
BIG_CUDA_PRED_BATCH = 4
BIG_CUDA_TRAIN_BATCH = 12
SMALL_CUDA_BATCH = 2
NUM_WORKERS = 4
PIN_CUDA_MEMORY = True

IM_SIZE_DIVISOR = 32      # images are padded to multiples of this
MODEL_INPUT_CHANNELS = 1  # greyscale
IMAGENET_MEAN = 0.449     # applied inside the slicer kernel
IMAGENET_STD = 0.226

DEFAULT_MIN_LR = 0.00075
LR_DIVISOR = 3
