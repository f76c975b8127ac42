"""Real-dependency pin of the read conditioning (scripts/STRique.py:590-595) for the bundled read.

What runs here is REAL library code wherever this container has it: `scipy.signal.medfilt`, numpy's median, and
`scipy.ndimage.grey_erosion / grey_dilation` (the C code scikit-image's grey morphology is a thin wrapper of).  The
only part restated is that wrapper -- scikit-image < 0.15 (`skimage/morphology/grey.py`, pinned by the reference's
requirements.txt:9) is not installable here; the four functions below follow its source: the even 1x8 structuring
element is zero-padded to 9 taps (`_shift_selem`), on the other side for the second pass of opening / closing,
and dilation hands scipy the reversed element.

Output: tests/golden/condition_pin.npz = uint8 codes after closing(opening(.)) of data/c9orf72.fast5, plus median and
MAD of the filtered signal.  The oracle's window restatement (oracle/reference_path.py: MORPH_WINDOWS) and the CUDA
conditioning kernel must reproduce it byte for byte (tests/test_condition_pin.py).

    python -m tests.golden.make_condition_pin
"""
import os
import sys

import numpy as np
import scipy.ndimage as ndi
import scipy.signal as sp

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


# ---- scikit-image 0.14 grey morphology wrapper (restated; the work is done by scipy.ndimage) ----------------------
def _shift_selem(selem, shift_x, shift_y):
    m, n = selem.shape
    if m % 2 == 0:
        extra_row = np.zeros((1, n), selem.dtype)
        selem = np.vstack((selem, extra_row)) if shift_x else np.vstack((extra_row, selem))
        m += 1
    if n % 2 == 0:
        extra_col = np.zeros((m, 1), selem.dtype)
        selem = np.hstack((selem, extra_col)) if shift_y else np.hstack((extra_col, selem))
    return selem


def erosion(image, selem, shift_x=False, shift_y=False):
    out = np.empty_like(image)
    ndi.grey_erosion(image, footprint=_shift_selem(np.array(selem), shift_x, shift_y), output=out)
    return out


def dilation(image, selem, shift_x=False, shift_y=False):
    selem = _shift_selem(np.array(selem), shift_x, shift_y)
    out = np.empty_like(image)
    ndi.grey_dilation(image, footprint=selem[::-1, ::-1], output=out)
    return out


def opening(image, selem):
    return dilation(erosion(image, selem), selem, shift_x=True, shift_y=True)


def closing(image, selem):
    return erosion(dilation(image, selem), selem, shift_x=True, shift_y=True)


def condition_u8(raw):
    """S.py:590-595 with the real scipy / numpy calls"""
    flt = sp.medfilt(raw, kernel_size=3)
    med = np.median(flt)
    mad = np.mean(np.absolute(np.subtract(flt, med)))                       # pore_model.MAD, S.py:142-143
    morph = (flt - med) / mad
    morph = np.clip(morph * 24 + 127, 0, 255).astype(np.dtype('uint8')).reshape((1, len(morph)))
    flt_selem = np.ones((1, 8), dtype=np.uint8)                             # skimage.morphology.rectangle(1, 8)
    morph = opening(morph, flt_selem)
    morph = closing(morph, flt_selem)[0]
    return flt, float(med), float(mad), morph


def main():
    from strique_b200 import fast5
    raw = fast5.read_raw_signal(os.path.join(ROOT, 'data', 'c9orf72.fast5'))
    flt, med, mad, u8 = condition_u8(raw)
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'condition_pin.npz'), u8=u8, flt_crc=np.int64(int(np.bitwise_xor.reduce(flt.astype(np.int64) * np.arange(1, len(flt) + 1)))),
                        median=med, mad=mad, n=len(raw), scipy=np.array(__import__('scipy').__version__))
    print('wrote condition_pin.npz:', len(u8), 'codes, median', med, 'MAD', mad)


if __name__ == '__main__':
    main()
