// glk/pointcloud_buffer.hpp STAND-IN (test infrastructure): the viewer is not part of the path
#pragma once
