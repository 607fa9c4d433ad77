/* Aborting definitions for the TH accessors declared in stub/TH/TH.h. */
#include <stdlib.h>
#include "TH/TH.h"
float* THFloatTensor_data(THFloatTensor* t) { (void)t; abort(); }
long THFloatTensor_size(THFloatTensor* t, int dim) { (void)t; (void)dim; abort(); }
THFloatStorage* THFloatTensor_storage(THFloatTensor* t) { (void)t; abort(); }
void THFloatStorage_fill(THFloatStorage* s, float v) { (void)s; (void)v; abort(); }
