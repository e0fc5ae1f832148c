// updater_kernels.cuh -- (placeholder, filled in by updater.cu)
#pragma once
