/* stub header: third-party dependency absent offline; the hot path does not use it. Written for the oracle shim build, not copied from anywhere. */
#pragma once
