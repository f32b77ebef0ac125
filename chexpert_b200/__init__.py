"""chexpert_b200 -- B200-native AAConv2d hot path of kamenbliznashki/chexpert."""
from .aaconv import AAConv2d, AAConvFunction  # noqa: F401
from .loss import BCEWithLogitsLoss, COMPETITION_INDEX  # noqa: F401
