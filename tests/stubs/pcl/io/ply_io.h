// TEST STUB (see registration/registration.h)
