"""Import alias: the package directory is `mpc-sensorlessao_b200/` (hyphen, as the project is
named); Python cannot import a hyphenated name, so this module loads that directory as the
package `mpc_sensorlessao_b200`."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "mpc-sensorlessao_b200")
_spec = _ilu.spec_from_file_location("mpc_sensorlessao_b200", _os.path.join(_pkg_dir, "__init__.py"),
                                     submodule_search_locations=[_pkg_dir])
_mod = _ilu.module_from_spec(_spec)
_sys.modules["mpc_sensorlessao_b200"] = _mod
_spec.loader.exec_module(_mod)
