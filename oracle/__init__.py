"""CPU oracle for the volume-segmantics prediction hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``volume_segmantics_b200/`` or
``volume_segmantics/`` may import this package; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / reference legs do.

What it restates (reference file:line, relative to /root/reference):

* ``vol_seg_2d_predictor.py:31-116``  prediction loop, softmax/argmax/gather,
  3-way and 12-way max-probability merges                -> predict_oracle.py
* ``datasets.py:120-142``, ``augmentations.py:30-65``      slice -> reflect-101
  centre pad to /32 -> /255 -> (x-0.449)/0.226            -> predict_oracle.py
* ``base_data_utils.py:125-138``      centre crop (banker's rounding), axis swap
* ``model_2d.py:10-57``               .pytorch dict -> model  -> smp_models.py

Parity status: **parity unpinned for network arithmetic**.  The reference
cannot be imported in this image (h5py, segmentation_models_pytorch,
albumentations are absent and not in the wheelhouse) and its own tests assert
only dtypes/shapes, never values.  The third-party
arithmetic is restated from the packages' published behaviour:
segmentation-models-pytorch ^0.2.1 (pyproject.toml:24), albumentations ^1.1.0
(pyproject.toml:19).  What *is* pinned: ``get_padded_dimension`` known answers
(tests/test_augmentations.py:6-10), the use of the very same
``cv2.copyMakeBorder`` / ``torchvision.center_crop`` / ``torchvision ResNet``
/ numpy ops the reference reaches, and dtype/shape contracts.
"""
