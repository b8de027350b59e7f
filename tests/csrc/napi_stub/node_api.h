// Declarations-only stand-in for <node_api.h> (node is not in the build image): just enough of the N-API surface that
// simple_zk_rollups_b200/ts/zkr_napi.cc uses, with the signatures of the real header (Node.js N-API v3), so that
// tests/test_abi.py can at least compile the addon source (g++ -fsyntax-only).  Nothing here is linked or run.
#pragma once
#include <stddef.h>
#include <stdint.h>
extern "C" {
typedef struct napi_env__* napi_env;
typedef struct napi_value__* napi_value;
typedef struct napi_callback_info__* napi_callback_info;
typedef struct napi_deferred__* napi_deferred;
typedef struct napi_async_work__* napi_async_work;
typedef enum { napi_ok, napi_invalid_arg, napi_generic_failure } napi_status;
typedef enum { napi_undefined, napi_null, napi_boolean, napi_number, napi_string, napi_symbol, napi_object, napi_function,
               napi_external, napi_bigint } napi_valuetype;
typedef enum { napi_int8_array, napi_uint8_array, napi_uint8_clamped_array, napi_int16_array, napi_uint16_array,
               napi_int32_array, napi_uint32_array, napi_float32_array, napi_float64_array } napi_typedarray_type;
typedef napi_value (*napi_callback)(napi_env env, napi_callback_info info);
typedef void (*napi_finalize)(napi_env env, void* finalize_data, void* finalize_hint);
typedef void (*napi_async_execute_callback)(napi_env env, void* data);
typedef void (*napi_async_complete_callback)(napi_env env, napi_status status, void* data);
#define NAPI_AUTO_LENGTH SIZE_MAX
napi_status napi_throw_error(napi_env env, const char* code, const char* msg);
napi_status napi_throw_type_error(napi_env env, const char* code, const char* msg);
napi_status napi_get_cb_info(napi_env env, napi_callback_info cbinfo, size_t* argc, napi_value* argv, napi_value* this_arg, void** data);
napi_status napi_get_arraybuffer_info(napi_env env, napi_value arraybuffer, void** data, size_t* byte_length);
napi_status napi_get_buffer_info(napi_env env, napi_value value, void** data, size_t* length);
napi_status napi_create_external(napi_env env, void* data, napi_finalize finalize_cb, void* finalize_hint, napi_value* result);
napi_status napi_get_value_external(napi_env env, napi_value value, void** result);
napi_status napi_get_typedarray_info(napi_env env, napi_value typedarray, napi_typedarray_type* type, size_t* length, void** data,
                                     napi_value* arraybuffer, size_t* byte_offset);
napi_status napi_get_boolean(napi_env env, bool value, napi_value* result);
napi_status napi_create_arraybuffer(napi_env env, size_t byte_length, void** data, napi_value* result);
napi_status napi_create_typedarray(napi_env env, napi_typedarray_type type, size_t length, napi_value arraybuffer, size_t byte_offset,
                                   napi_value* result);
napi_status napi_resolve_deferred(napi_env env, napi_deferred deferred, napi_value resolution);
napi_status napi_reject_deferred(napi_env env, napi_deferred deferred, napi_value rejection);
napi_status napi_create_string_utf8(napi_env env, const char* str, size_t length, napi_value* result);
napi_status napi_create_error(napi_env env, napi_value code, napi_value msg, napi_value* result);
napi_status napi_delete_async_work(napi_env env, napi_async_work work);
napi_status napi_typeof(napi_env env, napi_value value, napi_valuetype* result);
napi_status napi_create_promise(napi_env env, napi_deferred* deferred, napi_value* promise);
napi_status napi_create_async_work(napi_env env, napi_value async_resource, napi_value async_resource_name,
                                   napi_async_execute_callback execute, napi_async_complete_callback complete, void* data,
                                   napi_async_work* result);
napi_status napi_queue_async_work(napi_env env, napi_async_work work);
napi_status napi_create_function(napi_env env, const char* utf8name, size_t length, napi_callback cb, void* data, napi_value* result);
napi_status napi_set_named_property(napi_env env, napi_value object, const char* utf8name, napi_value value);
}
#define NAPI_MODULE(modname, regfunc) extern "C" napi_value napi_register_module_v1(napi_env env, napi_value exports) { return regfunc(env, exports); }
