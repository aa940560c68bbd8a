"""Drop-in mirror of the reference's ``miniplant`` package for the photon-tracing hot path
(reference ``miniplant/simulation_runner.py``, ``scene_creator.py``, ``utils.py`` and the two batch
drivers).  Same module names, function names, argument meaning and error behaviour; the tracing is
done by the sm_100a CUDA library behind ``include/lscpm.h``."""
