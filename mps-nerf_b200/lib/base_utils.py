"""IO helpers kept for API parity with the reference's ``lib/base_utils.py:6-10``."""
import pickle


def read_pickle(pkl_path):
    """Load a (python-2 era) SMPL pickle with latin1 decoding."""
    with open(pkl_path, "rb") as fh:
        return pickle.load(fh, encoding="latin1")
