"""Oracle restatement of the reference prediction loop (CPU, fp32).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows, statement by statement:
  volume_segmantics/model/operations/vol_seg_2d_predictor.py:31-116
  volume_segmantics/data/datasets.py:120-142        (slice -> pad -> normalise)
  volume_segmantics/data/augmentations.py:30-65     (pad sizes; A.PadIfNeeded
        defaults [ext]: centre position, cv2.BORDER_REFLECT_101,
        pad_top = int(p / 2.0), pad_bottom = p - pad_top)
  volume_segmantics/utilities/base_data_utils.py:125-138 (crop, axis rotation)
  volume_segmantics/utilities/config.py:35,41-42    (32, 0.449, 0.226)

The reference hard-requires CUDA (base_data_utils.py:107-108); the only
deliberate deviation is running the same ops on the CPU with a caller-chosen
batch size (the reference uses 4 on any >= 8 GB GPU, config.py:29-32).
"""
from __future__ import annotations

import math

import cv2
import numpy as np
import torch
import torchvision.transforms.functional as TVF

IM_SIZE_DIVISOR = 32  # config.py:35
IMAGENET_MEAN = 0.449  # config.py:41
IMAGENET_STD = 0.226  # config.py:42

AXIS_Z, AXIS_Y, AXIS_X = 0, 1, 2


def get_padded_dimension(dimension: int) -> int:
    """augmentations.py:30-44; known answers tests/test_augmentations.py:6-10."""
    if dimension % IM_SIZE_DIVISOR == 0:
        return dimension
    return (math.floor(dimension / IM_SIZE_DIVISOR) + 1) * IM_SIZE_DIVISOR


def rotate_array_to_axis(array: np.ndarray, axis: int) -> np.ndarray:
    """base_data_utils.py:132-138 (views, self-inverse)."""
    if axis == AXIS_Z:
        return array
    if axis == AXIS_Y:
        return array.swapaxes(0, 1)
    if axis == AXIS_X:
        return array.swapaxes(0, 2)
    raise ValueError(axis)


def pad_slice(image: np.ndarray) -> np.ndarray:
    """A.PadIfNeeded(min_height=pad32(H), min_width=pad32(W)) [ext] ->
    cv2.copyMakeBorder(..., BORDER_REFLECT_101) with the floor/ceil split."""
    rows, cols = image.shape[:2]
    ph = get_padded_dimension(rows) - rows
    pw = get_padded_dimension(cols) - cols
    top = int(ph / 2.0)
    left = int(pw / 2.0)
    if ph == 0 and pw == 0:
        return image
    # cv2 rejects some integer dtypes (e.g. int64) directly; albumentations
    # hands the array to cv2 unchanged, which converts int64 -> int32.
    src = image
    if src.dtype == np.int64:
        src = src.astype(np.int32)
    return cv2.copyMakeBorder(
        np.ascontiguousarray(src), top, ph - top, left, pw - left, cv2.BORDER_REFLECT_101
    )


def preprocess_slice(image: np.ndarray) -> np.ndarray:
    """datasets.py:120-142 -> float32 [Hp, Wp]."""
    image = pad_slice(image)
    if np.issubdtype(image.dtype, np.integer):
        image = image.astype(np.float32)
        image = image / 255
    image = image - IMAGENET_MEAN
    image = image / IMAGENET_STD
    return image


def crop_tensor_to_array(t: torch.Tensor, yx_dims) -> np.ndarray:
    """base_data_utils.py:125-129: torchvision center_crop (banker's rounding
    of the crop offset) then numpy."""
    return TVF.center_crop(t, list(yx_dims)).detach().numpy()


class OraclePredictor:
    """Mirror of VolSeg2dPredictor (vol_seg_2d_predictor.py:16-136) for a model
    object supplied by the caller."""

    def __init__(self, model: torch.nn.Module, num_labels: int, batch_size: int = 4):
        self.model = model.eval()
        self.num_labels = num_labels
        self.batch_size = batch_size

    # -- network on padded/normalised slices --------------------------------
    def logits_for_slices(self, slices: np.ndarray) -> torch.Tensor:
        """slices [S,H,W] any dtype -> logits fp32 [S,C,Hp,Wp]."""
        outs = []
        with torch.no_grad():
            for s0 in range(0, slices.shape[0], self.batch_size):
                batch = np.stack(
                    [preprocess_slice(slices[i]) for i in range(s0, min(s0 + self.batch_size, slices.shape[0]))]
                )
                x = torch.from_numpy(np.ascontiguousarray(batch, dtype=np.float32))[:, None]
                outs.append(self.model(x))
        return torch.cat(outs)

    # -- vol_seg_2d_predictor.py:31-65 ---------------------------------------
    def predict_single_axis(self, data_vol, output_probs=True, axis=AXIS_Z, return_full=False):
        data_vol = rotate_array_to_axis(data_vol, axis)
        yx_dims = list(data_vol.shape[1:])
        label_list, prob_list, full_list = [], [], []
        s_max = torch.nn.Softmax(dim=1)
        with torch.no_grad():
            for s0 in range(0, data_vol.shape[0], self.batch_size):
                batch = np.stack(
                    [preprocess_slice(data_vol[i]) for i in range(s0, min(s0 + self.batch_size, data_vol.shape[0]))]
                )
                x = torch.from_numpy(np.ascontiguousarray(batch, dtype=np.float32))[:, None]
                output = self.model(x)
                probs = s_max(output)
                if return_full:
                    full_list.append(crop_tensor_to_array(probs, yx_dims))
                labels = torch.argmax(probs, dim=1)
                labels = crop_tensor_to_array(labels, yx_dims)
                label_list.append(labels.astype(np.uint8))
                if output_probs:
                    idx = torch.argmax(probs, dim=1, keepdim=True)
                    p = torch.squeeze(torch.gather(probs, 1, idx), dim=1)
                    p = crop_tensor_to_array(p, yx_dims)
                    prob_list.append(p.astype(np.float16))
        labels = rotate_array_to_axis(np.concatenate(label_list), axis)
        probs = np.concatenate(prob_list) if prob_list else None
        if probs is not None:
            probs = rotate_array_to_axis(probs, axis)
        if return_full:
            return labels, probs, np.concatenate(full_list)  # [S,C,H,W] in slice space
        return labels, probs

    # -- :67-98 ---------------------------------------------------------------
    def predict_3_ways_max_probs(self, data_vol):
        shape = data_vol.shape
        label_c = np.empty((2, *shape), dtype=np.uint8)
        prob_c = np.empty((2, *shape), dtype=np.float16)
        label_c[0], prob_c[0] = self.predict_single_axis(data_vol, True, AXIS_Z)
        label_c[1], prob_c[1] = self.predict_single_axis(data_vol, True, AXIS_Y)
        merge_vols_in_mem(prob_c, label_c)
        label_c[1], prob_c[1] = self.predict_single_axis(data_vol, True, AXIS_X)
        merge_vols_in_mem(prob_c, label_c)
        return label_c[0], prob_c[0]

    # -- :100-116 ---------------------------------------------------------------
    def predict_12_ways_max_probs(self, data_vol):
        shape = data_vol.shape
        label_c = np.empty((2, *shape), dtype=np.uint8)
        prob_c = np.empty((2, *shape), dtype=np.float16)
        label_c[0], prob_c[0] = self.predict_3_ways_max_probs(data_vol)
        for k in range(1, 4):
            data_vol = np.rot90(data_vol)
            labels, probs = self.predict_3_ways_max_probs(data_vol)
            label_c[1] = np.rot90(labels, -k)
            prob_c[1] = np.rot90(probs, -k)
            merge_vols_in_mem(prob_c, label_c)
        return label_c[0], prob_c[0]

    # -- margin of a merged decision (test helper, not in the reference) ---------------
    def class_best_over_directions(self, data_vol, dirs):
        """[C,Z,Y,X] fp32: for every class the largest softmax probability any of the listed
        directions assigns to it at a voxel.  The top-2 gap of this array is the reference's
        margin between the label that wins the max-probability merge (:90-98) and the best
        competing label -- the quantity BASELINE.json's "every disagreement at a voxel whose
        reference top-2 margin is below tolerance" refers to for 3-way / 12-way results."""
        out = None
        for d in dirs:
            sl = np.ascontiguousarray(direction_slices(data_vol, d))
            _, _, full = self.predict_single_axis(sl, True, AXIS_Z, return_full=True)  # [S,C,H,W]
            vol_c = np.stack([direction_to_volume(full[:, c], d) for c in range(full.shape[1])])
            out = vol_c if out is None else np.maximum(out, vol_c)
        return out

    # -- :118-136 one-hot vote variants ------------------------------------------
    def predict_single_axis_to_one_hot(self, data_vol, axis=AXIS_Z):
        pred, _ = self.predict_single_axis(data_vol, axis=axis)
        return one_hot_encode_array(pred, self.num_labels)

    def predict_3_ways_one_hot(self, data_vol):
        out = self.predict_single_axis_to_one_hot(data_vol)
        out += self.predict_single_axis_to_one_hot(data_vol, AXIS_Y)
        out += self.predict_single_axis_to_one_hot(data_vol, AXIS_X)
        return out

    def predict_12_ways_one_hot(self, data_vol):
        out = self.predict_3_ways_one_hot(data_vol)
        for k in range(1, 4):
            data_vol = np.rot90(data_vol)
            out += np.rot90(self.predict_3_ways_one_hot(data_vol), -k, axes=(-3, -2))
        return out


def merge_vols_in_mem(prob_container: np.ndarray, label_container: np.ndarray) -> None:
    """vol_seg_2d_predictor.py:90-98 (in place on slot 0; fp16 compare, ties -> slot 0)."""
    idx = np.argmax(prob_container, axis=0)[np.newaxis]
    prob_container[0] = np.squeeze(np.take_along_axis(prob_container, idx, axis=0), axis=0)
    label_container[0] = np.squeeze(np.take_along_axis(label_container, idx, axis=0), axis=0)


def one_hot_encode_array(input_array: np.ndarray, num_labels: int) -> np.ndarray:
    """base_data_utils.py:141-147."""
    out = np.zeros((num_labels, input_array.size), dtype=np.uint8)
    out[input_array.ravel(), np.arange(input_array.size)] = 1
    out.shape = (num_labels,) + input_array.shape
    return out


# ---------------------------------------------------------------------------
# Injected-probability oracle: everything EXCEPT the network.  Used for the
# bit-exact slicing / merge criterion of BASELINE.json.
# ---------------------------------------------------------------------------

def direction_slices(data_vol: np.ndarray, d: int) -> np.ndarray:
    """The (S,H,W) stack of images the reference feeds the network for
    direction d = 3k + a (k rot90 steps in the Z-Y plane, a = Z/Y/X axis):
    vol_seg_2d_predictor.py:108 + :34."""
    k, a = divmod(d, 3)
    return rotate_array_to_axis(np.rot90(data_vol, k), a)


def direction_to_volume(arr_sHW: np.ndarray, d: int) -> np.ndarray:
    """Inverse: a per-slice result stack of direction d back to the original
    (Z,Y,X) orientation: vol_seg_2d_predictor.py:61-64 then :110-111."""
    k, a = divmod(d, 3)
    return np.rot90(rotate_array_to_axis(arr_sHW, a), -k)


def slicer_oracle(data_vol: np.ndarray, d: int) -> np.ndarray:
    """Padded + normalised fp32 input stack [S,Hp,Wp] for direction d."""
    sl = direction_slices(data_vol, d)
    return np.stack([preprocess_slice(sl[i]) for i in range(sl.shape[0])]).astype(np.float32)


def crop_offsets(h: int, w: int):
    """(crop_top, crop_left) torchvision.center_crop uses from the padded image."""
    hp, wp = get_padded_dimension(h), get_padded_dimension(w)
    return int(round((hp - h) / 2.0)), int(round((wp - w) / 2.0))


def merge_injected_oracle(shape_zyx, dirs, probs_by_dir, labels_by_dir):
    """First-max merge of per-direction (prob fp32 [S,H,W], label uint8 [S,H,W])
    stacks given in *slice space* of each direction, exactly as the reference
    would fold them: fp16 cast (predictor.py:58), inverse rotation (:61-64,
    :110-111), nested 3-way/12-way first-wins folds (:67-116) which compose to
    a flat earliest-direction-wins max over the listed directions in order."""
    label_c = np.zeros((2, *shape_zyx), dtype=np.uint8)
    prob_c = np.zeros((2, *shape_zyx), dtype=np.float16)
    first = True
    for d in dirs:
        lab = direction_to_volume(labels_by_dir[d].astype(np.uint8), d)
        prb = direction_to_volume(probs_by_dir[d].astype(np.float16), d)
        if first:
            label_c[0], prob_c[0] = lab, prb
            first = False
        else:
            label_c[1], prob_c[1] = lab, prb
            merge_vols_in_mem(prob_c, label_c)
    return label_c[0].copy(), prob_c[0].copy()


def clip_to_uint8_oracle(data: np.ndarray, data_mean: float, st_dev_factor: float) -> np.ndarray:
    """base_data_utils.py:243-287 restated (numpy, float64): clip to mean +- k*sigma, rescale,
    truncate.  Works on a copy; the reference mutates float input in place."""
    data = np.array(data, copy=True)
    data_st_dev = np.nanstd(data)
    lower_bound = data_mean - (data_st_dev * st_dev_factor)
    upper_bound = data_mean + (data_st_dev * st_dev_factor)
    if np.isnan(data).any():
        data = np.nan_to_num(data, copy=False, nan=data_mean)
    if np.issubdtype(data.dtype, np.integer):
        data = data.astype(float)
    data = np.clip(data, lower_bound, upper_bound, out=data)
    data = np.subtract(data, lower_bound, out=data)
    data = np.divide(data, (upper_bound - lower_bound), out=data)
    data = np.clip(data, 0.0, 1.0, out=data)
    data = np.multiply(data, 255, out=data)
    return data.astype(np.uint8)
