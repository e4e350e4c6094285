// TEST INFRASTRUCTURE: forwards to the TensorFlow stand-in (see ../../../tf_stub.h).
#pragma once
#include "../../../tf_stub.h"
