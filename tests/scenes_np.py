"""Scene recipes live in the package (playrho_b200/scenes_np.py: bench.py and __graft_entry__.smoke() use them too)."""
from playrho_b200.scenes_np import *  # noqa: F401,F403
from playrho_b200.scenes_np import SceneBuilder, GRAVITY  # noqa: F401
