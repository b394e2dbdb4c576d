"""CPU oracle for the downstream estimators.  TEST INFRASTRUCTURE ONLY.

Value-level numpy restatement of ``simpleDS/utils.py``: ``bootstrap_array`` (171-209),
``weighted_average`` (569-645, in ``simpleds_oracle``) and ``fold_along_delay`` (651-796).
PINNED: ``tests/golden/make_estimators_golden.py`` executes the reference's unmodified
``utils.py`` (stub astropy namespace, array-function-aware Quantity) on seeded inputs and
``tests/test_estimators.py`` requires these functions to reproduce
``tests/golden/estimators_golden.npz`` bit for bit, next to the closed-form literals of the
reference's own tests (``tests/test_utils.py:563-766``).  Only ``tests/`` may import this.
"""
import numpy as np

from .simpleds_oracle import weighted_average


def bootstrap_array(array, nboot=100, axis=0):
    """utils.py:171-209; draws from numpy's global generator like the reference."""
    if axis >= len(np.shape(array)):
        raise ValueError(
            "Specified axis must be shorter than the length "
            "of input array.\n"
            "axis value was {0} but array has {1} dimensions".format(axis, len(np.shape(array)))
        )
    sample_inds = np.random.choice(array.shape[axis], size=(array.shape[axis], nboot), replace=True)
    return np.take(array, sample_inds, axis=axis)


def fold_along_delay(delays, array, uncertainty, weights=None, axis=-1):
    """utils.py:651-796 on plain arrays (delays in any one time unit)."""
    delays = np.array(delays, copy=True)
    array = np.array(array, copy=True)
    uncertainty = np.array(uncertainty, copy=True)
    if array.shape[axis] != len(delays):  # utils.py:698-707
        raise ValueError("Input array must have same length as the delays along the specified axis.")
    if len(delays) % 2 != 0 and np.abs(delays).min() != 0:  # utils.py:708-714
        raise ValueError("Input delays must have either a delay=0 bin as the central value or have an even size.")
    if not array.shape == uncertainty.shape:  # utils.py:717-722
        raise ValueError("Input array and uncertainties must have the same shape.")
    if weights is None:  # utils.py:724-728
        if not array.imag.any():
            weights = np.ones_like(array)
        else:
            weights = np.ones_like(array) * (1 + 1j)

    if np.logical_and(np.abs(delays).min() == 0, delays.size % 2):  # utils.py:730-757
        split_index = np.argmin(np.abs(delays), axis=axis)
        split_inds = [split_index, split_index + 1]
        neg_vals, zero_bin, pos_vals = np.split(array, split_inds, axis=axis)
        neg_vals = np.flip(neg_vals, axis=axis)
        pos_vals = np.concatenate([zero_bin, pos_vals], axis=axis)
        neg_vals = np.concatenate([zero_bin, neg_vals], axis=axis)
        neg_errors, zero_errors, pos_errors = np.split(uncertainty, split_inds, axis=axis)
        zero_errors = zero_errors * np.sqrt(2)  # utils.py:737 (in place on the split view's copy)
        neg_errors = np.flip(neg_errors, axis=axis)
        pos_errors = np.concatenate([zero_errors, pos_errors], axis=axis)
        neg_errors = np.concatenate([zero_errors, neg_errors], axis=axis)
        neg_weights, zero_weights, pos_weights = np.split(weights, split_inds, axis=axis)
        neg_weights = np.flip(neg_weights, axis=axis)
        pos_weights = np.concatenate([zero_weights, pos_weights], axis=axis)
        neg_weights = np.concatenate([zero_weights, neg_weights], axis=axis)
    else:  # utils.py:759-770
        min_val_bool = np.abs(delays) == np.amin(np.abs(delays), axis=axis)
        split_index = np.where(np.logical_and(delays >= 0, min_val_bool))
        split_inds = [np.squeeze(split_index)]
        neg_vals, pos_vals = np.split(array, split_inds, axis=axis)
        neg_vals = np.flip(neg_vals, axis=axis)
        neg_errors, pos_errors = np.split(uncertainty, split_inds, axis=axis)
        neg_errors = np.flip(neg_errors, axis=axis)
        neg_weights, pos_weights = np.split(weights, split_inds, axis=axis)
        neg_weights = np.flip(neg_weights, axis=axis)

    _array = np.stack([pos_vals, neg_vals], axis=0)  # utils.py:772-774
    _errors = np.stack([pos_errors, neg_errors], axis=0)
    _weights = np.stack([pos_weights, neg_weights], axis=0)

    if not _array.imag.any():  # utils.py:776-779
        return weighted_average(_array.real, _errors.real, weights=_weights, axis=0)
    weight_check = _weights.imag.any()  # utils.py:781-787
    if not weight_check:
        try:
            _weights.imag = np.ones_like(_weights.real)
        except TypeError:  # real weights: the reference falls to SINGLE precision here
            _weights = _weights.astype(np.complex64)
            _weights.imag = np.ones_like(_weights.real)
    out_array_real, out_errors_real = weighted_average(_array.real, _errors.real, weights=_weights.real, axis=0)
    out_array_imag, out_errors_imag = weighted_average(_array.imag, _errors.imag, weights=_weights.imag, axis=0)
    return out_array_real + 1j * out_array_imag, out_errors_real + 1j * out_errors_imag
