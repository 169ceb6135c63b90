"""CPU: the algorithmic-FLOP model behind roofline.achieved reproduces the survey's examples."""
import pytest


@pytest.mark.parametrize("cls,flops", [((0, 0, 0, 0), 128), ((1, 0, 1, 0), 572), ((1, 1, 1, 1), 8699), ((2, 0, 2, 0), 2885),
                                       ((2, 1, 2, 1), 44882), ((2, 2, 2, 2), 311387)])
def test_model_flops_match_survey_examples(cls, flops):
    from unomol_b200 import capi
    assert capi.lib.unomol_b200_model_flops(*cls) == flops
