"""Import alias: the package lives in `spark-sched-sim_b200/` (a name Python cannot import
directly); `import spark_sched_sim_b200` loads that directory as a regular package."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "spark-sched-sim_b200")
_spec = _ilu.spec_from_file_location(
    "spark_sched_sim_b200", _os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir]
)
_mod = _ilu.module_from_spec(_spec)
_sys.modules["spark_sched_sim_b200"] = _mod
_spec.loader.exec_module(_mod)
