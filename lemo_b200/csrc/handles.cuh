// Opaque C-ABI handle wrappers (shared by the translation units that implement extern "C" entry points).
#pragma once
#include "body.cuh"
#include "vposer.cuh"
#include "conv.cuh"
struct LemoModel { lemo::Model* m; };
struct LemoBody { lemo::BodyCtx* c; };
struct LemoVPoser { lemo::VPoser* v; };
struct LemoConvNet { lemo::ConvNet* n; float* dx_planes; };
