// TEST INFRASTRUCTURE: stands in for the TensorFlow header of the same path (see tests/tf_mock/tf_mock.h)
#include "tf_mock.h"
