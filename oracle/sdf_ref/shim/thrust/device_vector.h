// TEST INFRASTRUCTURE (oracle/sdf_ref): the reference includes <thrust/device_vector.h> but uses nothing of it.
#pragma once
