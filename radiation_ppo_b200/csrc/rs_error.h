// internal: thread-local last-error text shared by the translation units of libradsearch_b200
#pragma once
int rs_set_error(const char *msg);   // always returns -1
