"""biseqt_b200 -- B200-native replacement for the pairwise-alignment hot path
of biseqt (``biseqt.pw`` / ``pwlib``).  See DESIGN.md."""
from .sequence import Alphabet, Sequence  # noqa: F401
