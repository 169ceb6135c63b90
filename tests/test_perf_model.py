"""CPU: the algorithmic-FLOP model behind roofline.achieved reproduces the survey's examples."""
import pytest


@pytest.mark.parametrize("cls,flops", [((0, 0, 0, 0), 128), ((1, 0, 1, 0), 572), ((1, 1, 1, 1), 8699), ((2, 0, 2, 0), 2885),
                                       ((2, 1, 2, 1), 44882), ((2, 2, 2, 2), 311387)])
def test_model_flops_match_survey_examples(cls, flops):
    try:
        from unomol_b200 import capi     # needs the built library and the CUDA runtime libraries it links (no device needed)
    except (ImportError, OSError) as e:
        pytest.skip("libunomol_b200.so cannot be loaded here: %s" % e)
    assert capi.lib.unomol_b200_model_flops(*cls) == flops
