// The engine behind tt_engine_*: per-GPU weight replicas, streams and workspaces; pages are
// sharded over GPUs by one host worker thread per device, results gathered on the host.
#pragma once
#include "tuatara_c.h"

struct tt_engine;  // defined in engine.cpp
