"""Starts the reference's stock server entry point (`infinistore.server.main`, from
baseline/_ref) with ONE accommodation for the sandbox: `prevent_oom()` writes -1000 to
/proc/<pid>/oom_score_adj, which an unprivileged container refuses (PermissionError) and the
reference does not catch - the process would die right after its pool and listener came up.
The function is unrelated to the data path; everything else is the reference's own code."""
import sys

import infinistore.server as server


def _prevent_oom_if_allowed():
    try:
        server_prevent_oom()
    except (PermissionError, OSError):
        pass


server_prevent_oom = server.prevent_oom
server.prevent_oom = _prevent_oom_if_allowed

if __name__ == "__main__":
    sys.argv[0] = "infinistore.server"
    server.main()
