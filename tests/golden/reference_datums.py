"""Literal golden datums held by the reference's own tests, with the values they decode to.

Sources (file:line in /root/reference):
  G1  ruhvro/src/deserialize.rs:244 (schema :187-242)  — reference asserts 8 columns x 4 rows (:248-249)
  G2  ruhvro/src/deserialize.rs:303 (schema :254-301)  — reference asserts 5 columns x 1 row (:307-308)
  G3-G5 ruhvro/src/lib.rs:165-167 (schema :65-161)     — reference asserts "decodes without error" (:174)
The reference asserts only shapes; the VALUES below were decoded by hand/script from the hex
 and are asserted by tests/test_oracle_golden.py against both oracles and,
on the GPU, against the product.
"""
import json

G1_SCHEMA = json.dumps({
    "type": "record", "name": "UserData", "namespace": "com.example",
    "fields": [
        {"name": "userId", "type": "string"},
        {"name": "age", "type": "int"},
        {"name": "fullName", "type": {"type": "record", "name": "FullName", "fields": [
            {"name": "firstName", "type": "string"}, {"name": "lastName", "type": "string"}]}},
        {"name": "email", "type": ["null", "string"], "default": None},
        {"name": "phoneNumbers", "type": {"type": "array", "items": "string"}},
        {"name": "isPremiumMember", "type": "boolean"},
        {"name": "favoriteItems", "type": {"type": "map", "values": "int"}},
        {"name": "registrationDate", "type": {"type": "long", "logicalType": "timestamp-millis"}},
    ]})
G1_HEX = ("4834346437643065662d613264662d343833652d393261312d313532333830366164656334380a4c696e64610857617265022c6c"
          "696e646173636f7474406578616d706c652e6e6574062628323636293734302d31323737783031313432283030312d3935392d38"
          "39342d36353030783739392a3030312d3339362d3831392d363830307830303139000006044d72100866696e640e10617070726f"
          "6163680c00c0f691c7c35f")
G1_ROW = {
    "userId": "44d7d0ef-a2df-483e-92a1-1523806adec4", "age": 28,
    "fullName": {"firstName": "Linda", "lastName": "Ware"},
    "email": "lindascott@example.net",
    "phoneNumbers": ["(266)740-1277x01142", "001-959-894-6500x799", "001-396-819-6800x0019"],
    "isPremiumMember": False,
    "favoriteItems": [("Mr", 8), ("find", 7), ("approach", 6)],
    "registrationDate_ms": 1641154756000,
}

G2_SCHEMA = json.dumps({
    "type": "record", "name": "User", "namespace": "com.example",
    "fields": [
        {"name": "firstName", "type": "string"},
        {"name": "lastName", "type": "string"},
        {"name": "age", "type": "int"},
        {"name": "addresses", "type": {"type": "array", "items": {"type": "record", "name": "Address", "fields": [
            {"name": "street", "type": "string"}, {"name": "city", "type": "string"}, {"name": "zipCode", "type": "string"}]}}},
        {"name": "email", "type": ["null", "string"], "default": "null"},
    ]})
G2_HEX = ("084a6f686e06446f653c041431323320456c6d20537412536f6d6577686572650a313233343514343536204f616b20537410416e"
          "7977686572650a36373839300002286a6f686e2e646f65406578616d706c652e636f6d")
G2_ROW = {
    "firstName": "John", "lastName": "Doe", "age": 30,
    "addresses": [{"street": "123 Elm St", "city": "Somewhere", "zipCode": "12345"},
                  {"street": "456 Oak St", "city": "Anywhere", "zipCode": "67890"}],
    "email": "john.doe@example.com",
}

G345_SCHEMA = json.dumps({
    "type": "record", "name": "User",
    "fields": [
        {"name": "name", "type": ["null", "string"], "default": None},
        {"name": "age", "type": ["null", "int"], "default": None},
        {"name": "emails", "type": {"type": "array", "items": "string"}},
        {"name": "address", "type": ["null", {"type": "record", "name": "Address", "fields": [
            {"name": "street", "type": "string"}, {"name": "city", "type": "string"}, {"name": "zipcode", "type": "string"}]}],
         "default": None},
        {"name": "phone_numbers", "type": {"type": "map", "values": "string"}},
        {"name": "preferences", "type": ["null", {"type": "record", "name": "Preferences", "fields": [
            {"name": "contact_method", "type": ["null", "string"], "default": None},
            {"name": "newsletter", "type": "boolean"}]}], "default": None},
        {"name": "status", "type": ["string", "int", "boolean"]},
    ]})
G3_HEX = ("0000062e74686f6d61736b6172656e406578616d706c652e6e657422616c6f7765406578616d706c652e6f72672664617669643738"
          "406578616d706c652e636f6d0000060a636865636b203030312d3233372d3438302d353133341065766964656e6365262b312d37"
          "32352d3336362d39323133783730300a6d616a6f722428393734293537302d3032313178333534350002020a656d61696c00020a"
          "7374616666")
G4_HEX = ("0218416d616e646120456c6c6973023804246e6361736579406578616d706c652e636f6d307374657761727474796c6572406578"
          "616d706c652e6e6574000230393532323120436861726c657320547261666669637761791c5a616368617279626f726f7567680a"
          "303433343300000202")
G5_HEX = ("021a417564726579204261726e65730000000408726963682628373136293338322d363937327837383437320e746f6e69676874"
          "243536302e3736352e3230363378373831313500020000000866726565")
G345_ROWS = [
    {"name": None, "age": None,
     "emails": ["thomaskaren@example.net", "alowe@example.org", "david78@example.com"],
     "address": None,
     "phone_numbers": [("check", "001-237-480-5134"), ("evidence", "+1-725-366-9213x700"), ("major", "(974)570-0211x3545")],
     "preferences": {"contact_method": "email", "newsletter": False},
     "status": 5, "status_type_id": 1},
    {"name": "Amanda Ellis", "age": 28,
     "emails": ["ncasey@example.com", "stewarttyler@example.net"],
     "address": {"street": "95221 Charles Trafficway", "city": "Zacharyborough", "zipcode": "04343"},
     "phone_numbers": [],
     "preferences": None,
     "status": 1, "status_type_id": 1},
    {"name": "Audrey Barnes", "age": None,
     "emails": [],
     "address": None,
     "phone_numbers": [("rich", "(716)382-6972x78472"), ("tonight", "560.765.2063x78115")],
     "preferences": {"contact_method": None, "newsletter": False},
     "status": "free", "status_type_id": 0},
]
