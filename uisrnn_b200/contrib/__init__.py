"""Community-contributed helpers kept for API compatibility with `uisrnn.contrib` of the
reference (`/root/reference/uisrnn/contrib/`).  Not on the predict()/fit() path."""
from . import contrib_template
from . import range_search_crp_alpha
