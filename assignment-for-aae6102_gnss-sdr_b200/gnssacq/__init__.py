"""gnssacq -- host-side mirror of the reference's acquisition interface over libgnssacq.so."""
from . import api  # noqa: F401  (raises ImportError if libgnssacq.so has not been built)
from .api import Config, Result, Stats, Channel, LoopParams, TrackRecord, Searcher, GnssAcqError, make_config, default_config, version  # noqa: F401
from .acquisition import acquisition, config_from_structs, get_searcher, release_all  # noqa: F401
from .params import initParameters  # noqa: F401
from .matfile import save_acquired, load_acquired, acquired_filename, cached_acquisition  # noqa: F401
from .tracking import trackingCT, trackParameters, cn0_estimates, bit_edge_index  # noqa: F401
