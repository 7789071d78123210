#include <assert.h>
#include "fea_model.h"

void fea_model_init(fea_model_ptr self, model_type type) {
  assert(type == MODEL_A5 || type == MODEL_COMPRESSIBLE_NEOHOOKEAN);
  self->model = type;
  if (self->parameters_count < 2) self->parameters_count = 2;
}

const char *fea_model_name(model_type type) {
  return type == MODEL_A5 ? "A5" : "COMPRESSIBLE_NEOHOOKEAN";
}
