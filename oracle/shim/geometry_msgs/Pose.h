// TEST INFRASTRUCTURE ONLY: see ../ros_msgs_shim.h
#include "../ros_msgs_shim.h"
