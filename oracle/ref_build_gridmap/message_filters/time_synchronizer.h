// stand-in (TEST INFRASTRUCTURE): see subscriber.h
#pragma once
#include <message_filters/subscriber.h>
