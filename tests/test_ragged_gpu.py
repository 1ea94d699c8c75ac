"""Ragged / odd configurations of the training step against the oracle: uneven branch splits (n_no_noise = int(B p),
train_ae.py:304), odd batch sizes, mask ratios whose len_keep = int(L (1 - r)) (ae.py:11) is not a multiple of any tile
size, masked DiT (mask on the noise branch only), very high mask ratios, one-sample branches."""
import pytest

from tests import util as U

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kw", [
    dict(batch=6, no_noise_prob=0.25),                                   # 5 noised + 1 clean
    dict(batch=5, no_noise_prob=0.5),                                    # 3 + 2 (int(2.5) = 2)
    dict(batch=3, no_noise_prob=0.9),                                    # 1 + 2
    dict(batch=4, mask_ratio=0.3, mask_ratio_no_noise=0.55),             # len_keep 179 / 115: S_e = 183 / 119
    dict(batch=4, mask_ratio=0.5, no_noise_prob=0.0),                    # masked DiT: one branch, masked
    dict(batch=4, mask_ratio=0.0, mask_ratio_no_noise=0.95),             # unmasked noise branch (S_e = 260) + 12 kept tokens
    dict(batch=4, mask_ratio=0.99, mask_ratio_no_noise=0.99),            # 2 kept tokens per sample
    dict(batch=2, adaln=False, mask_ratio=0.3, mask_ratio_no_noise=0.9), # cond-token blocks, odd lengths
], ids=lambda kw: ",".join(f"{k}={v}" for k, v in kw.items()))
def test_step_parity_ragged(kw):
  U.run_step_parity(variant="S/4", steps=1, depth=2, dec_depth=1, **kw)
